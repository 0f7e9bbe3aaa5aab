"""Scaling sweep (BASELINE.json configs[4]): synthetic cfg4-shaped scenes (5 000 tags, 25 % visibility) from ~1 M to
~200 M corner observations on 1 / 2 / 4 / 8 GPUs, the CPU restatement's figure beside every point.
One JSON line per size on rank 0's stdout.  The problem of a point is FIXED and sharded over the ranks by
keyframe range (strong scaling, like bench.py).

  python tools/sweep.py [M ...]                                             (1 GPU; M = millions of observations)
  python -m torch.distributed.run --nproc-per-node N ... tools/sweep.py [M ...]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload generator + CPU arm)


def main():
    import torch
    import torch.distributed as dist
    from robot_camera_calibration_b200.problem import BAProblem
    sizes = [float(x) for x in sys.argv[1:]] or [1, 2, 5, 10, 20, 50]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def red(x, op):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    obs_per_view = 4 * 0.25 * 5000 * 0.95
    for M in sizes:
        scale = M * 1e6 / obs_per_view / bench.CONFIGS[4][1]
        t0 = time.time()
        s, desc, views = bench.workload(4, rank, world, scale)
        t_gen = time.time() - t0
        gp = BAProblem.from_scene(s, device=local, eliminate="views")
        if world > 1:
            ids = [BAProblem.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            gp.comm_init(ids[0], rank, world)
        for _ in range(3):
            gp.linearize(want_cost=False)
        barrier()
        gp.profile_reset(); gp.profile_enable(True)
        steps = 5
        for _ in range(steps):
            gp.flush_l2(); gp.linearize(want_cost=False)
        barrier()
        lin = {k: v[0] / steps for k, v in gp.profile().items() if v[0] > 0}
        lin_ms = red(sum(lin.values()), dist.ReduceOp.MAX if world > 1 else None)

        def lm():
            gp.linearize(want_cost=False); gp.schur(1e4); gp.solve_step(); gp.candidate_cost()
        lm()
        barrier()
        gp.profile_reset()
        n_lm = 2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(n_lm):
            lm()
        gp.synchronize()
        barrier()
        lm_ms = red((time.perf_counter() - t1) * 1e3 / n_lm, dist.ReduceOp.MAX if world > 1 else None)
        stages = {k: v[0] / n_lm for k, v in gp.profile().items() if v[0] > 0}
        obs = red(s.n_observations, dist.ReduceOp.SUM if world > 1 else None)
        n_red = int(gp.dims.n_reduced)
        gp.close()
        barrier()
        if rank == 0:
            # CPU figure beside the point: the restatement on (at most) the first 1 250 keyframes of the problem.
            # Only in the 1-GPU run: under torchrun the other ranks spin in the barrier and would steal the cores;
            # the figure does not depend on N (same problem), so the N > 1 lines refer to the N = 1 line.
            cpu = {"value": None, "cores": None, "sample": "see the n_gpus = 1 line of the same size"}
            if world == 1:
                chunk = bench.CONFIGS[4][4]
                hi = max(chunk, min(1250, views) // chunk * chunk)
                cpu = bench.cpu_baseline(4, scale, [(0, hi)], f"keyframes 0..{hi - 1} of the point", seconds=2.0)
            print(json.dumps({"workload": desc, "n_gpus": world, "observations": int(obs), "views": views,
                              "n_reduced": n_red, "scene_generation_s": round(t_gen, 1),
                              "linearize_ms": lin_ms, "obs_per_s": obs / (lin_ms * 1e-3),
                              "lm_iter_ms": lm_ms, "lm_obs_per_s": obs / (lm_ms * 1e-3), "lm_stage_ms_rank0": stages,
                              "cpu_obs_per_s": cpu["value"], "cpu_cores": cpu["cores"], "cpu_sample": cpu["sample"][:60]}),
                  flush=True)
        barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
