"""CPU: the __host__ __device__ model functions of csrc/model.cuh (the arithmetic every
kernel runs) compiled with g++ through tests/host_harness/model_harness.cpp and checked
against the complex-step oracle -- catches math errors without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import ba_oracle as O
from helpers import to_oracle
from robot_camera_calibration_b200.scenes import make_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DP = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = tmp_path_factory.mktemp("harness") / "model_harness.so"
    subprocess.run(["/usr/bin/g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-o", str(so),
                    os.path.join(ROOT, "tests", "host_harness", "model_harness.cpp")], check=True)
    return ctypes.CDLL(str(so))


@pytest.mark.parametrize("model,ncam", [("single", 1), ("rig", 3)])
def test_analytic_jacobian_matches_complex_step(harness, model, ncam):
    s = make_scene(20, 14, 0.5, n_cam=ncam, model=model, seed=5)
    s.views[3, 0:3] = [1e-9, -2e-9, 3e-10]          # series branches of the Rodrigues coefficients
    s.markers[5, 0:3] = [1e-4, 2e-4, -1e-4]
    s.views[7, 0:3] *= 0.01
    p = to_oracle(s)
    r, Jb = O.residuals(p), O.jacobian_blocks_cs(p)
    P = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(DP)
    worst = 0.0
    for b in range(s.n_blocks):
        c = s.cam_idx[b]
        sh = np.concatenate([s.intr[c], s.dist[c]])
        r8, jv, jm, js, jx, d4 = np.zeros(8), np.zeros(48), np.zeros(48), np.zeros(72), np.zeros(48), np.zeros(4)
        harness.model_eval_block(int(model == "rig"), P(s.views[s.view_idx[b]]), P(s.markers[s.marker_idx[b]]),
                                 P(s.ext[c]), P(sh), ctypes.c_double(s.sizes[s.marker_idx[b]]), P(s.pixels[b]),
                                 r8.ctypes.data_as(DP), jv.ctypes.data_as(DP), jm.ctypes.data_as(DP),
                                 js.ctypes.data_as(DP), jx.ctypes.data_as(DP), d4.ctypes.data_as(DP))
        rel = lambda a, ref: np.linalg.norm(a - ref) / max(np.linalg.norm(ref), 1e-300)
        assert np.abs(r8 - r[b]).max() <= 1e-9 * max(1.0, np.abs(r[b]).max())
        errs = [rel(jv.reshape(8, 6), Jb["view"][b]), rel(jm.reshape(8, 6), Jb["marker"][b]),
                rel(js.reshape(8, 9)[:, :4], Jb["intr"][b]), rel(js.reshape(8, 9)[:, 4:], Jb["dist"][b])]
        if model == "rig":
            errs.append(rel(jx.reshape(8, 6), Jb["ext"][b]))
        worst = max(worst, max(errs))
    assert worst < 1e-12


def test_expand_pose_matches_opencv(harness):
    import cv2
    rng = np.random.default_rng(3)
    for r in [rng.normal(0, 1, 3), np.zeros(3), np.array([1e-8, 0, 0]), rng.normal(0, 0.05, 3), np.array([0, 3.0, 0.5])]:
        p6 = np.concatenate([r, [1.0, 2.0, 3.0]])
        out = np.zeros(24)
        harness.model_expand_pose(p6.ctypes.data_as(DP), out.ctypes.data_as(DP))
        R, _ = cv2.Rodrigues(r.reshape(3, 1))
        assert np.abs(out[0:9].reshape(3, 3) - R).max() < 1e-14
        assert np.allclose(out[18:21], [1.0, 2.0, 3.0])
        # right Jacobian: d(R(r) p)/dr = -R [p]x Jr  checked by complex step
        pvec = np.array([0.3, -0.2, 0.9])
        Jr = out[9:18].reshape(3, 3)
        num = np.zeros((3, 3))
        for k in range(3):
            rc = r.astype(complex); rc[k] += 1e-30j
            num[:, k] = (O.rodrigues(rc) @ pvec).imag / 1e-30
        px = np.array([[0, -pvec[2], pvec[1]], [pvec[2], 0, -pvec[0]], [-pvec[1], pvec[0], 0]])
        assert np.abs(-R @ px @ Jr - num).max() < 1e-13
