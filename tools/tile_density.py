"""Host-side analysis for the Schur SYRK: how densely do the tags a keyframe sees fill T-block
column tiles of the kept set, with the scene's own (random) tag numbering and with the tags renumbered along a Morton
curve of their positions?  `util` = useful block pairs / block pairs of the visited tiles taken as dense -- the share
of an MMA-tile (or masked register) formulation's work that would be useful.
usage: tile_density.py [cfg] [scale]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from robot_camera_calibration_b200.scenes import config_scene

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 4
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
s = config_scene(cfg, scale=scale, blocked=True)
nv, nm = len(s.views), len(s.markers)
pos = s.truth['markers'][:, 3:6]
c = pos - pos.mean(0)
_, _, vt = np.linalg.svd(c, full_matrices=False)
xy = c @ vt[:2].T                                         # coordinates in the wall's principal plane


def morton(xy, bits=8):
    q = ((xy - xy.min(0)) / (np.ptp(xy, 0) + 1e-9) * (2 ** bits - 1)).astype(np.int64)
    code = np.zeros(len(q), np.int64)
    for b in range(bits):
        code |= ((q[:, 0] >> b) & 1) << (2 * b)
        code |= ((q[:, 1] >> b) & 1) << (2 * b + 1)
    return code


order = np.argsort(morton(xy), kind='stable')
rank = np.empty(nm, int)
rank[order] = np.arange(nm)
out = {"scene": s.name, "views": nv, "tags": nm, "blocks_per_view": s.n_blocks / nv, "rows": []}
for name, lab in (('scene numbering', np.arange(nm)), ('Morton order', rank)):
    mi = lab[s.marker_idx]
    for T in (2, 4, 8, 16, 32):
        useful = dense = 0
        fill = []
        for v in range(nv):
            m = np.unique(mi[s.view_idx == v])
            t = np.unique(m // T)
            useful += len(m) * (len(m) + 1) / 2
            dense += (len(t) * T) * (len(t) * T + 1) / 2
            fill.append(len(m) / (len(t) * T))
        out["rows"].append({"numbering": name, "tile_blocks": T, "util": round(useful / dense, 3),
                            "mean_fill_of_visited_tiles": round(float(np.mean(fill)), 3)})
print(json.dumps(out, indent=1))
