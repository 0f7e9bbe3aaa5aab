"""The C++/OpenMP "Ceres-equivalent" restatement (oracle/cpu_restatement.cpp)
against the numpy complex-step oracle -- pins the timed CPU baseline."""
import numpy as np
import pytest

import ba_oracle as O
from cpu_baseline import CpuBA
from helpers import max_block_rel, oracle_blocks, oracle_reduced, rel_fro, to_oracle
from robot_camera_calibration_b200.scenes import make_scene

CASES = {"single": dict(n_cam=1, model="single", seed=21), "rig": dict(n_cam=3, model="rig", seed=22)}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("elim", ["views", "markers"])
def test_restatement_matches_oracle(name, elim):
    s = make_scene(10, 16, 0.7, **CASES[name])
    p = to_oracle(s)
    cpu = CpuBA(s, eliminate=elim)
    cost, fail = cpu.linearize()
    assert fail == 0
    out = cpu.get()
    r, Jb = O.residuals(p), O.jacobian_blocks_cs(p)
    assert max_block_rel(out["residuals"], r) < 1e-9
    for k in Jb:
        assert max_block_rel(out[k], Jb[k]) < 1e-9, k
    ob = oracle_blocks(p, elim == "views")
    assert abs(cost - ob["cost"]) <= 1e-12 * ob["cost"]
    for k in ("Hee", "ge", "Hes", "Hff", "gf", "Hfs", "W"):
        assert max_block_rel(out[k], ob[k], floor=1e-6 * np.abs(ob[k]).max()) < 1e-9, k
    assert rel_fro(out["Hss"], ob["Hss"]) < 1e-9
    assert rel_fro(out["gs"], ob["gs"]) < 1e-9
    # Schur complement and one LM step
    S, b, f_index, H, g, d2 = oracle_reduced(p, elim == "views", 1e4)
    Sc, bc, bad = cpu.schur(1e4)
    assert bad == 0
    assert rel_fro(Sc, S) < 1e-9
    assert rel_fro(bc, b) < 1e-9
    dE, dF, _ = cpu.lm_iteration(1e4)
    delta, mcc, c, gmax = O.lm_step(p, 1e4)
    o_view, o_marker, o_shared, n = p.offsets()
    dv, dm = delta[o_view:o_marker].reshape(-1, 6), delta[o_marker:o_shared].reshape(-1, 6)
    de, df = (dv, dm) if elim == "views" else (dm, dv)
    assert rel_fro(dE, de) < 1e-6
    assert rel_fro(dF, np.concatenate([df.ravel(), delta[o_shared:]])) < 1e-6
