"""Under torchrun (N ranks, one GPU each): the N-GPU LM solve must reach the same
parameters as a single-GPU solve of the whole problem.  Rank 0 prints a JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from robot_camera_calibration_b200.dist import DistributedBA
from robot_camera_calibration_b200.problem import BAProblem
from robot_camera_calibration_b200.scenes import make_scene

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = {}
for name, kw in (("single", dict(n_markers=40, n_views=120, visibility=0.5, seed=41)),
                 ("rig", dict(n_markers=30, n_views=60, visibility=0.5, n_cam=2, model="rig", seed=42)),
                 ("single_dist", dict(n_markers=70, n_views=140, visibility=0.4, seed=43)),
                 ("rig_dist", dict(n_markers=50, n_views=80, visibility=0.5, n_cam=2, model="rig", seed=44))):
    kw = dict(kw)
    # *_dist: the distributed hand-written Cholesky (block columns over the ranks, panels broadcast) forced on a
    # system of 3-4 panels; the others take the automatic choice (replicated cuSOLVER at this size)
    if name.endswith("_dist"):
        os.environ["RCC_CHOLESKY"] = "dist"
        os.environ["RCC_SYRK_GROUPS"] = "3"         # and several bands in the overlapped reduction
    else:
        os.environ.pop("RCC_CHOLESKY", None)
        os.environ.pop("RCC_SYRK_GROUPS", None)
    scene = make_scene(kw.pop("n_markers"), kw.pop("n_views"), kw.pop("visibility"), **kw)
    opts = dict(max_iterations=30, function_tolerance=1e-14, gradient_tolerance=1e-12, parameter_tolerance=1e-13)
    dba = DistributedBA(scene, device=local)
    summ = dba.solve(**opts)
    views, markers, intr, dst = dba.gather_parameters()
    dba.close()
    if rank == 0:
        with BAProblem.from_scene(scene, device=local) as gp:
            s1 = gp.solve(**opts)
            v1, m1 = gp.get_view_poses(), gp.get_marker_poses()
            i1, d1 = gp.get_intrinsics()
        out[name] = dict(iters=(summ["iterations"], s1["iterations"]), cost=(summ["final_cost"], s1["final_cost"]),
                         dviews=float(np.abs(views - v1).max()), dmarkers=float(np.abs(markers - m1).max()),
                         dintr=float(np.abs(intr - i1).max()), ddist=float(np.abs(dst - d1).max()),
                         allreduce_ms=summ["allreduce_ms"])
    dist.barrier()
if os.environ.get("RCC_CHECK_BIG") == "1":
    # one LM step of a rig problem whose reduced system (n = 12 060, cfg3's tags and cameras, 10 % of its body poses)
    # is large enough for the automatic choice of the distributed Cholesky and the packed all-reduce
    from robot_camera_calibration_b200.scenes import config_scene
    os.environ.pop("RCC_CHOLESKY", None)
    os.environ.pop("RCC_SYRK_GROUPS", None)
    scene = config_scene(3, scale=0.1, blocked=True)
    dba = DistributedBA(scene, device=local, eliminate="views")
    p = dba.problem
    p.linearize(); p.schur(1e4); p.solve_step()
    st = p.step()
    cand = p.candidate_cost()
    dba.close()
    if rank == 0:
        os.environ["RCC_CHOLESKY"] = "cusolver"
        with BAProblem.from_scene(scene, device=local, eliminate="views") as gp:
            gp.linearize(); gp.schur(1e4); gp.solve_step()
            s1 = gp.step()
            c1 = gp.candidate_cost()
        dF = np.concatenate([st["d_f"].ravel(), st["d_shared"]]); d1 = np.concatenate([s1["d_f"].ravel(), s1["d_shared"]])
        out["cfg3_one_step"] = dict(n_reduced=int(6 * len(scene.markers) + 15 * len(scene.intr)),
                                    delta_F_rel=float(np.linalg.norm(dF - d1) / np.linalg.norm(d1)),
                                    candidate_cost=(cand, c1), iters=(1, 1), cost=(cand, c1),
                                    dviews=0.0, dmarkers=float(np.abs(st["d_f"] - s1["d_f"]).max()), ddist=0.0)
    dist.barrier()
if rank == 0:
    ok = all(v["dviews"] < 1e-8 and v["dmarkers"] < 1e-8 and v["ddist"] < 1e-8 and
             abs(v["cost"][0] - v["cost"][1]) <= 1e-10 * v["cost"][1] for v in out.values())
    print(json.dumps({"world": world, "ok": ok, **out}))
    if not ok:
        sys.exit(1)
dist.destroy_process_group()
