"""Small end-to-end pass over every kernel (single + rig, both elimination directions, robust loss, update_pixels)
for compute-sanitizer:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from robot_camera_calibration_b200.problem import BAProblem
from robot_camera_calibration_b200.scenes import make_scene

for kw, elim in ((dict(), "views"), (dict(), "markers"), (dict(n_cam=2, model="rig"), "views")):
    s = make_scene(70, 45, 0.6, seed=71, **kw)
    with BAProblem.from_scene(s, eliminate=elim) as gp:
        gp.evaluate()
        gp.evaluate(want_jacobians=False)
        gp.linearize()
        gp.update_pixels(np.ascontiguousarray(s.pixels + 0.1))
        gp.linearize(); gp.schur(1e4); gp.solve_step(); gp.candidate_cost()
        gp.set_loss("huber", 1.0)
        summ = gp.solve(max_iterations=3)
        print(elim, kw.get("model", "single"), summ["initial_cost"], "->", summ["final_cost"], flush=True)
print("done")
