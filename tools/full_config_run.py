"""Full-size run of one BASELINE.json config on one GPU: scene generation (host), index build,
linearise timing, LM solve to the noise floor.  Prints one JSON line.
usage: full_config_run.py CFG [SCALE] [LM_ITERS]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from robot_camera_calibration_b200.problem import BAProblem
from robot_camera_calibration_b200.scenes import config_scene

cfg = int(sys.argv[1]); scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
lm_iters = int(sys.argv[3]) if len(sys.argv) > 3 else 12
t0 = time.time(); s = config_scene(cfg, scale=scale, blocked=True); t_gen = time.time() - t0
t0 = time.time(); gp = BAProblem.from_scene(s); t_setup = time.time() - t0
d = gp.dims
for _ in range(2):
    gp.linearize(want_cost=False)
gp.profile_reset(); gp.profile_enable(True)
steps = 5
for _ in range(steps):
    gp.flush_l2(); gp.linearize(want_cost=False)
gp.synchronize()
pr = {k: v[0] / steps for k, v in gp.profile().items() if v[0] > 0}
lin_ms = sum(pr.values())
gp.profile_enable(False)
# one untimed LM step: cuSOLVER / NCCL one-time initialisation (~1.4 s) is not an iteration cost
gp.linearize(want_cost=False); gp.schur(1e4); gp.solve_step(); gp.candidate_cost()
t0 = time.time()
summ = gp.solve(max_iterations=lm_iters, function_tolerance=1e-9)
t_solve = time.time() - t0
n_res = 8 * s.n_blocks
n_par = 6 * (len(s.views) + len(s.markers) - 1) + len(s.intr) * (15 if s.model == "rig" else 9)
floor = 0.5 * 0.3 ** 2 * (n_res - n_par)
markers = gp.get_marker_poses()
out = {"config": cfg, "scale": scale, "model": s.model, "n_views": len(s.views), "n_markers": len(s.markers),
       "n_cameras": len(s.intr), "blocks": s.n_blocks, "observations": s.n_observations,
       "eliminated": "views" if d.eliminated_is_view else "markers", "n_reduced": d.n_reduced,
       "scene_generation_s": round(t_gen, 1), "host_setup_s": round(t_setup, 1),
       "linearize_ms": lin_ms, "linearize_stage_ms": pr, "obs_per_s": s.n_observations / (lin_ms * 1e-3),
       "lm": {k: summ[k] for k in ("iterations", "accepted", "termination", "initial_cost", "final_cost", "total_ms",
                                   "linearize_ms", "schur_ms", "solve_ms", "backsub_ms", "cost_ms")},
       "s_per_lm_iter": summ["total_ms"] * 1e-3 / max(1, summ["iterations"]),
       "final_cost_over_noise_floor": summ["final_cost"] / floor,
       "marker_t_err_mean_m": float(np.abs(markers[:, 3:] - s.truth["markers"][:, 3:]).mean()),
       "solve_wall_s": round(t_solve, 1)}
print(json.dumps(out))
