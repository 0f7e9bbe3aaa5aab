"""CPU: the oracle against the committed golden vectors (OpenCV analytic
derivatives) and against itself (A vs B), plus its LM on a small scene."""
import numpy as np
import pytest

import ba_oracle as O
from helpers import GOLDEN, load_golden, max_block_rel, oracle_blocks, oracle_reduced, rel_fro, to_oracle
from robot_camera_calibration_b200.scenes import make_scene


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_a_matches_opencv_golden(name):
    s, g = load_golden(name)
    p = to_oracle(s)
    assert np.abs(O.residuals(p) - g["residuals"]).max() < 1e-9        # pixels
    Jb = O.jacobian_blocks_cs(p)
    for k in Jb:
        assert max_block_rel(Jb[k], g[f"jac_{k}"]) < 1e-10, k   # OpenCV itself loses digits near r = 0


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_normal_equations_match_golden(name):
    s, g = load_golden(name)
    p = to_oracle(s)
    for elim in ("views", "markers"):
        ob = oracle_blocks(p, elim == "views")
        for k in ("Hee", "ge", "Hes", "Hff", "gf", "Hfs", "Hss", "gs", "W"):
            assert rel_fro(ob[k], g[f"{elim}_{k}"]) < 1e-12, (elim, k)
        S, b, *_ = oracle_reduced(p, elim == "views", float(g["radius"]))
        assert rel_fro(S, g[f"{elim}_S"]) < 1e-10
        assert rel_fro(b, g[f"{elim}_b"]) < 1e-10
    delta, mcc, cost, gmax = O.lm_step(p, float(g["radius"]))
    assert rel_fro(delta, g["delta"]) < 1e-8
    assert abs(cost - float(g["cost"])) < 1e-12 * cost


def test_schur_of_both_directions_gives_the_same_step():
    """Eliminating views or markers must give the same damped Gauss-Newton step."""
    s = make_scene(8, 10, 0.8, seed=5)
    p = to_oracle(s)
    H, g, _ = O.normal_equations(p)
    cm = p.const_mask()
    Hd = H + np.diag(O.lm_diagonal(H, 1e4))
    Hd, gm = O.masked_system(Hd, g, cm)
    full = -np.linalg.solve(Hd, gm)
    o_view, o_marker, o_shared, n = p.offsets()
    for e_slice in (slice(o_view, o_marker), slice(o_marker, o_shared)):
        f_index = np.array([i for i in range(n) if not (e_slice.start <= i < e_slice.stop)])
        S, b = O.schur_reduce(Hd, gm, e_slice, f_index)
        dF = -np.linalg.solve(S, b)
        assert rel_fro(dF, full[f_index]) < 1e-8


def test_lm_converges_to_truth_without_noise():
    s = make_scene(8, 12, 0.9, seed=7, pixel_noise=0.0)
    p = to_oracle(s)
    hist = O.lm_solve(p, max_iters=40)
    assert hist[-1] < 1e-12 * hist[0]
    assert np.abs(p.views - s.truth["views"]).max() < 1e-6
    assert np.abs(p.markers - s.truth["markers"]).max() < 1e-6
    assert np.abs(p.dist - s.truth["dist"]).max() < 1e-6


def test_rodrigues_matches_opencv():
    import cv2
    rng = np.random.default_rng(0)
    for r in [rng.normal(0, 1, 3), np.array([1e-9, 0, 0]), np.zeros(3), rng.normal(0, 1e-4, 3)]:
        R, _ = cv2.Rodrigues(r.reshape(3, 1))
        assert np.abs(O.rodrigues(r) - R).max() < 1e-14
