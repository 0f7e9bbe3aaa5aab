"""Times the materialised evaluation K1 on cfg2 (kernel-tuning helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import workload
from robot_camera_calibration_b200.problem import BAProblem
s, _, _ = workload(2, 0, 1, 1.0)
gp = BAProblem.from_scene(s)
for _ in range(3):
    gp.evaluate_device(True)
gp.profile_reset(); gp.profile_enable(True)
for _ in range(10):
    gp.flush_l2(); gp.evaluate_device(True)
ms = gp.profile()["evaluate"][0] / 10
print("K1 ms", round(ms, 4), "GB/s", round(1480 * s.n_blocks / ms / 1e6, 1), "frac", round(1480 * s.n_blocks / ms / 1e6 / 6538, 4))
