"""CPU: the BASELINE-size (block-sparse) restatement of the oracle used by the full-size GPU parity tests
must equal the dense oracle (dense J, dense J^T J, dense Schur complement) on scenes small enough for both."""
import numpy as np
import pytest

import ba_oracle as O
from helpers import (SparseSchurOracle, max_block_rel, oracle_blocks, oracle_blocks_sparse, oracle_reduced, rel_fro,
                     to_oracle)
from robot_camera_calibration_b200.scenes import make_scene


@pytest.mark.parametrize("model,elim_view,loss", [("single", True, None), ("single", False, None),
                                                   ("rig", True, None), ("rig", False, "huber")])
def test_sparse_oracle_equals_dense_oracle(model, elim_view, loss):
    kw = dict(n_cam=2, model="rig") if model == "rig" else {}
    s = make_scene(9, 11, 0.8, seed=71, **kw)
    s.const_views[2] = True
    p = to_oracle(s)
    if loss:
        p.loss, p.loss_scale = loss, 0.7
    dense = oracle_blocks(p, elim_view)
    sparse = oracle_blocks_sparse(p, elim_view, slab=37)        # several slabs
    assert abs(dense["cost"] - sparse["cost"]) <= 1e-13 * dense["cost"]
    for k in ("Hee", "ge", "Hes", "Hff", "gf", "Hfs", "W"):
        assert max_block_rel(sparse[k], dense[k], floor=1e-9 * np.abs(dense[k]).max()) < 1e-12, k
    assert rel_fro(sparse["Hss"], dense["Hss"]) < 1e-12 and rel_fro(sparse["gs"], dense["gs"]) < 1e-12
    S, b, *_ = oracle_reduced(p, elim_view, 1e4)
    so = SparseSchurOracle(p, sparse, elim_view, 1e4)
    n_f, ns = sparse["n_f"], sparse["n_shared"]
    scale = np.abs(S).max()
    for f in range(n_f):
        for g in range(n_f):
            assert np.abs(so.block(f, g) - S[6 * f:6 * f + 6, 6 * g:6 * g + 6]).max() < 1e-11 * scale
        bd = so.border(f)
        assert np.abs(bd[:, :ns] - S[6 * f:6 * f + 6, 6 * n_f:]).max() < 1e-11 * scale
        assert np.abs(bd[:, ns] - b[6 * f:6 * f + 6]).max() < 1e-11 * np.abs(b).max()
    c = so.corner()
    assert np.abs(c[:, :ns] - S[6 * n_f:, 6 * n_f:]).max() < 1e-11 * scale
    assert np.abs(c[:, ns] - b[6 * n_f:]).max() < 1e-11 * np.abs(b).max()
