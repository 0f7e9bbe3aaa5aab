"""Generates tests/golden/*.npz -- small fixed problems with inputs and expected
outputs.  Run here (build container), commit the .npz files; nothing on the GPU
box regenerates them.

What pins what: the reference ships no optimiser, no tests and no
golden vectors, so the expected values come from
  * oracle B = OpenCV's own analytic derivatives (cv2.projectPoints /
    composeRT / Rodrigues; cv2 version recorded in the file) for residuals and
    Jacobians -- the library the reference calls at camera_pose.cpp:163;
  * oracle A (numpy closed form + complex step) for the normal equations, the
    Schur complement and one LM step (dense linear algebra on A's Jacobian).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import cv2  # noqa: E402

import ba_oracle as O  # noqa: E402
from helpers import oracle_blocks, oracle_reduced, to_oracle  # noqa: E402
from robot_camera_calibration_b200.scenes import make_scene  # noqa: E402

CASES = {
    "single_small": dict(n_markers=6, n_views=8, visibility=0.8, seed=101),
    "rig_small": dict(n_markers=8, n_views=6, visibility=0.7, n_cam=2, model="rig", seed=102),
    "single_intpix": dict(n_markers=5, n_views=7, visibility=0.9, seed=103, round_pixels=True),
}


def main():
    for name, kw in CASES.items():
        kw = dict(kw)
        s = make_scene(kw.pop("n_markers"), kw.pop("n_views"), kw.pop("visibility"), **kw)
        if name == "single_small":
            s.markers[2, 0:3] = [1e-7, -2e-7, 5e-8]     # near-identity rotation: series branch
        p = to_oracle(s)
        N = s.n_blocks
        res = np.empty((N, 8))
        J = {k: np.empty((N, 8, n)) for k, n in O.local_param_names(s.model)}
        for b in range(N):
            c = s.cam_idx[b]
            uv, Jb = O.jacobian_block_cv2(s.model, s.intr[c], s.dist[c], s.ext[c], s.views[s.view_idx[b]],
                                          s.markers[s.marker_idx[b]], s.sizes[s.marker_idx[b]])
            res[b] = uv - s.pixels[b]
            for k in J:
                J[k][b] = Jb[k]
        out = dict(model=s.model, intr=s.intr, dist=s.dist, ext=s.ext, views=s.views, markers=s.markers,
                   sizes=s.sizes, view_idx=s.view_idx, marker_idx=s.marker_idx, cam_idx=s.cam_idx, pixels=s.pixels,
                   const_views=s.const_views, const_markers=s.const_markers, const_intr=s.const_intr,
                   const_dist=s.const_dist, const_ext=s.const_ext, cv2_version=cv2.__version__,
                   residuals=res, **{f"jac_{k}": v for k, v in J.items()})
        for elim in ("views", "markers"):
            ob = oracle_blocks(p, elim == "views")
            for k in ("Hee", "ge", "Hes", "Hff", "gf", "Hfs", "Hss", "gs", "W"):
                out[f"{elim}_{k}"] = ob[k]
            S, b_, f_index, H, g, d2 = oracle_reduced(p, elim == "views", 1e4)
            out[f"{elim}_S"], out[f"{elim}_b"] = S, b_
        delta, mcc, cost, gmax = O.lm_step(p, 1e4)
        out.update(cost=cost, delta=delta, mcc=mcc, radius=1e4)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, "blocks", N, "cost", cost)


if __name__ == "__main__":
    main()
