"""Wall-clock breakdown of one end-to-end step through the C ABI (host buffers): which call costs what."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import workload
from robot_camera_calibration_b200.problem import BAProblem, _dp

scene, desc, _ = workload(2, 0, 1, 1.0)
gp = BAProblem.from_scene(scene, eliminate="views")
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
h_pix, h_views, h_markers = pin(scene.pixels), pin(scene.views.copy()), pin(scene.markers.copy())
h_intr, h_dist = pin(scene.intr), pin(scene.dist)
d = gp.dims
h_ge, h_gf, h_gs = pin(np.zeros((d.n_e, 6))), pin(np.zeros((d.n_f, 6))), pin(np.zeros(d.n_shared))
acc = {}
def t(name, fn):
    t0 = time.perf_counter(); r = fn(); gp.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0; return r
N = 20
for it in range(N + 3):
    if it == 3:
        acc.clear()
    t("set_view_poses", lambda: gp.set_view_poses(h_views))
    t("set_marker_poses", lambda: gp.set_marker_poses(h_markers))
    t("set_intrinsics", lambda: gp.set_intrinsics(h_intr, h_dist))
    t0 = time.perf_counter()
    gp.update_pixels(h_pix)                    # asynchronous: pieces land on the side stream ...
    gp.linearize(want_cost=True)               # ... and the E pass of a piece starts when it has landed
    acc["update_pixels+linearize+cost"] = acc.get("update_pixels+linearize+cost", 0.0) + time.perf_counter() - t0
    t("get_normal_blocks", lambda: gp._check(gp.lib.rcc_ba_get_normal_blocks(gp.h, None, _dp(h_ge), None, None, _dp(h_gf), None, None, _dp(h_gs), None)))
print(json.dumps({k: round(v / N * 1e3, 4) for k, v in acc.items()}), "ms per step; total", round(sum(acc.values()) / N * 1e3, 4))
