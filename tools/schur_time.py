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
if os.environ.get("RCC_MORTON") == "1":
    # experiment: relabel the tags along a Morton curve of their positions (what would the SYRK gain from an
    # ordering in which every keyframe's tags are contiguous runs?)
    import numpy as np
    t = scene.markers[:, 3:6]
    q = ((t - t.min(0)) / np.maximum(t.max(0) - t.min(0), 1e-12) * 1023).astype(np.int64)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF0000FF
        v = (v | (v << 8)) & 0x0300F00F00F00F
        v = (v | (v << 4)) & 0x030C30C30C30C3
        v = (v | (v << 2)) & 0x09249249249249
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    code[0] = -1                                   # the world tag stays first
    order = np.argsort(code, kind="stable")        # new -> old
    inv = np.empty_like(order); inv[order] = np.arange(len(order))
    scene.markers, scene.sizes = scene.markers[order], scene.sizes[order]
    scene.const_markers = scene.const_markers[order]
    scene.marker_idx = inv[scene.marker_idx].astype(np.int32)
    desc += " [tags relabelled in Morton order]"
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
