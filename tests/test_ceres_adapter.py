"""The Ceres front end as code: include/rcc_ceres_adapter.h compiled with g++ against
tests/ceres_mock/ceres/ceres.h -- a stand-in that declares exactly the published Ceres signatures
(CostFunction::Evaluate, SizedCostFunction<8,4,5,6,6>, EvaluationCallback::PrepareForEvaluation,
Problem::AddResidualBlock / SetParameterBlockConstant) -- and linked against librcc_ba.so.

CPU: the adapter and its driver compile and link.  GPU: the driver builds the problem through AddResidualBlock,
evaluates it the way Ceres's evaluator walks residual blocks, and the residuals / per-block row-major Jacobians
that come out of Evaluate() equal the oracle's."""
import os
import subprocess

import numpy as np
import pytest

import ba_oracle as O
from helpers import max_block_rel, to_oracle
from robot_camera_calibration_b200 import _lib as L
from robot_camera_calibration_b200.scenes import make_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "adapter_driver")
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "tests", "ceres_mock"),
                    "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "ceres_mock", "adapter_driver.cpp"),
                    "-o", exe, "-L" + libdir, "-lrcc_ba", "-Wl,-rpath," + libdir], check=True)
    return exe


def test_adapter_compiles_against_the_published_ceres_signatures(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
def test_adapter_evaluate_matches_the_oracle(tmp_path):
    exe = _build(tmp_path)
    s = make_scene(14, 23, 0.8, seed=81)
    scene_bin, out_bin = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    with open(scene_bin, "wb") as f:
        np.array([len(s.views), len(s.markers), s.n_blocks], np.int64).tofile(f)
        for a in (s.intr[0], s.dist[0], s.views, s.markers, s.sizes):
            np.ascontiguousarray(a, np.float64).tofile(f)
        np.ascontiguousarray(s.view_idx, np.int32).tofile(f)
        np.ascontiguousarray(s.marker_idx, np.int32).tofile(f)
        np.ascontiguousarray(s.pixels, np.float64).tofile(f)
    r = subprocess.run([exe, scene_bin, out_bin], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = np.fromfile(out_bin, np.float64)
    meta = raw[:4].view(np.int64)
    n = s.n_blocks
    assert meta[0] == 1 and meta[1] == 8 * n and meta[2] == n * 8 * 21
    assert meta[3] == 3       # residual-only, + Jacobians at the same point, new point: one GPU evaluation each
    o = 4
    r0, o = raw[o:o + 8 * n].reshape(n, 8), o + 8 * n
    r1, o = raw[o:o + 8 * n].reshape(n, 8), o + 8 * n
    J1, o = raw[o:o + n * 168].reshape(n, 168), o + n * 168
    r2, o = raw[o:o + 8 * n].reshape(n, 8), o + 8 * n
    J2 = raw[o:o + n * 168].reshape(n, 168)

    def check(p, res, J):
        want_r, Jb = O.residuals(p), O.jacobian_blocks_cs(p)
        assert max_block_rel(res, want_r) < 1e-9
        split = {"intr": J[:, 0:32].reshape(n, 8, 4), "dist": J[:, 32:72].reshape(n, 8, 5),
                 "view": J[:, 72:120].reshape(n, 8, 6), "marker": J[:, 120:168].reshape(n, 8, 6)}
        world = s.marker_idx == 0
        for k in ("intr", "dist", "view"):
            assert max_block_rel(split[k], Jb[k]) < 1e-9, k
        # the world tag is constant: Ceres passes jacobians[3] == NULL for its blocks, nothing is written there
        assert max_block_rel(split["marker"][~world], Jb["marker"][~world]) < 1e-9
        assert np.all(split["marker"][world] == 0.0)

    p = to_oracle(s)
    assert np.array_equal(r0, r1)
    check(p, r1, J1)
    p.views = p.views + 1e-3
    p.intr = p.intr.copy()
    p.intr[0, 0] *= 1.001
    check(p, r2, J2)
