"""Times the Schur stages (prep, SYRK + border, shared) on one workload (kernel-tuning helper).
usage: schur_time.py [cfg] [scale] [iters]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from robot_camera_calibration_b200.problem import BAProblem

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
scene, desc, _ = workload(cfg, 0, 1, scale)
gp = BAProblem.from_scene(scene, eliminate="views")
gp.linearize(want_cost=False)
for _ in range(2):
    gp.schur(1e4)
gp.profile_reset(); gp.profile_enable(True)
for _ in range(iters):
    gp.flush_l2()
    gp.schur(1e4)
gp.synchronize()
pr = {k: round(v[0] / iters * 1e3, 1) for k, v in gp.profile().items() if v[0] > 0}
d = gp.dims
print(json.dumps({"desc": desc, "blocks": scene.n_blocks, "n_reduced": d.n_reduced, "us": pr}))
