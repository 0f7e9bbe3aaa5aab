"""Size sweep on one GPU (BASELINE.json configs[4]): K2 throughput and one LM iteration on cfg4-shaped scenes
(5 000 tags, 25 % visibility) from ~1 M to ~48 M corner observations.  One JSON line per size.
usage: sweep.py [scale ...]   (scale = fraction of the 10 000 keyframes)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from robot_camera_calibration_b200.problem import BAProblem

scales = [float(x) for x in sys.argv[1:]] or [0.02, 0.08, 0.33]
for sc in scales:
    t0 = time.time(); s, desc = workload(4, 0, sc); t_gen = time.time() - t0
    gp = BAProblem.from_scene(s, eliminate="views")
    for _ in range(3):
        gp.linearize(want_cost=False)
    gp.profile_reset(); gp.profile_enable(True)
    steps = 5
    for _ in range(steps):
        gp.flush_l2(); gp.linearize(want_cost=False)
    gp.synchronize()
    lin = {k: v[0] / steps for k, v in gp.profile().items() if v[0] > 0}
    lin_ms = sum(lin.values())
    gp.linearize(want_cost=False); gp.schur(1e4); gp.solve_step(); gp.candidate_cost()   # warm cuSOLVER
    gp.profile_reset()
    n_lm = 2
    for _ in range(n_lm):
        gp.linearize(want_cost=False); gp.schur(1e4); gp.solve_step(); gp.candidate_cost()
    gp.synchronize()
    lm = {k: v[0] / n_lm for k, v in gp.profile().items() if v[0] > 0}
    d = gp.dims
    print(json.dumps({"workload": desc, "observations": s.n_observations, "blocks": s.n_blocks,
                      "n_reduced": d.n_reduced, "scene_generation_s": round(t_gen, 1),
                      "linearize_ms": lin_ms, "obs_per_s": s.n_observations / (lin_ms * 1e-3),
                      "lm_iter_ms": sum(lm.values()), "lm_stage_ms": lm}), flush=True)
    gp.close()
