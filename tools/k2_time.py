"""Times the linearisation stages on one workload (kernel-tuning helper).
env: RCC_BA_LIB (variant library), RCC_CHUNK.  usage: k2_time.py [cfg] [scale] [steps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bench import workload
from robot_camera_calibration_b200.problem import BAProblem

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
lm = int(sys.argv[4]) if len(sys.argv) > 4 else 0
scene, desc, _ = workload(cfg, 0, 1, scale)
gp = BAProblem.from_scene(scene, eliminate="views")
for _ in range(3):
    gp.linearize(want_cost=False)
gp.profile_reset(); gp.profile_enable(True)
for _ in range(steps):
    gp.flush_l2()
    gp.linearize(want_cost=False)
gp.synchronize()
pr = gp.profile()
out = {k: round(v[0] / steps * 1e3, 1) for k, v in pr.items() if v[0] > 0}
tot = sum(out.values())
res = {"lib": os.path.basename(os.environ.get("RCC_BA_LIB", "default")), "chunk": os.environ.get("RCC_CHUNK", "64"),
       "blocks": scene.n_blocks, "us": out, "obs_per_s": 4 * scene.n_blocks / (tot * 1e-6)}
if lm:
    gp.profile_reset()
    for _ in range(lm):
        gp.linearize(want_cost=False); gp.schur(1e4); gp.solve_step(); gp.candidate_cost()
    gp.synchronize()
    res["lm_us"] = {k: round(v[0] / lm * 1e3, 1) for k, v in gp.profile().items() if v[0] > 0}
    pass
print(json.dumps(res))
