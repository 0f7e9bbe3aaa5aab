"""Shared test helpers: scene -> oracle problem, oracle block extraction."""
import numpy as np

import ba_oracle as O


def to_oracle(s):
    return O.OracleProblem(s.model, s.intr, s.dist, s.ext, s.views, s.markers, s.sizes, s.view_idx,
                           s.marker_idx, s.cam_idx, s.pixels, s.const_views, s.const_markers, s.const_intr,
                           s.const_dist, s.const_ext if s.model == "rig" else None)


def rel_fro(a, b):
    """||a-b||_F / ||b||_F (per-block parity measure)."""
    a, b = np.asarray(a), np.asarray(b)
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (den if den > 0 else 1.0)


def max_block_rel(a, b, floor=0.0):
    """max over leading index of the per-block relative Frobenius error.  Blocks
    whose reference norm is below `floor` are compared absolutely against floor."""
    a = np.asarray(a).reshape(len(a), -1)
    b = np.asarray(b).reshape(len(b), -1)
    num = np.linalg.norm(a - b, axis=1)
    den = np.maximum(np.linalg.norm(b, axis=1), floor if floor > 0 else 1e-300)
    return float((num / den).max()) if len(a) else 0.0


def oracle_blocks(p, elim_view):
    """Normal-equation blocks in the GPU library's layout from the dense oracle."""
    H, g, cost = O.normal_equations(p)
    o_view, o_marker, o_shared, n = p.offsets()
    nv, nm = len(p.views), len(p.markers)
    e_off, n_e = (o_view, nv) if elim_view else (o_marker, nm)
    f_off, n_f = (o_marker, nm) if elim_view else (o_view, nv)
    ns = n - o_shared
    blk = lambda off, i: slice(off + 6 * i, off + 6 * i + 6)
    out = {
        "Hee": np.stack([H[blk(e_off, i), blk(e_off, i)] for i in range(n_e)]),
        "ge": g[e_off:e_off + 6 * n_e].reshape(n_e, 6),
        "Hes": np.stack([H[blk(e_off, i), o_shared:] for i in range(n_e)]),
        "Hff": np.stack([H[blk(f_off, i), blk(f_off, i)] for i in range(n_f)]),
        "gf": g[f_off:f_off + 6 * n_f].reshape(n_f, 6),
        "Hfs": np.stack([H[blk(f_off, i), o_shared:] for i in range(n_f)]),
        "Hss": H[o_shared:, o_shared:],
        "gs": g[o_shared:],
        "cost": cost, "H": H, "g": g,
        "e_off": e_off, "n_e": n_e, "f_off": f_off, "n_f": n_f, "n_shared": ns,
    }
    # per-observation cross blocks W = J_e^T J_f
    Jb = O.jacobian_blocks_cs(p)
    Je, Jf = (Jb["view"], Jb["marker"]) if elim_view else (Jb["marker"], Jb["view"])
    w, _ = O.robust(p, O.residuals(p))
    out["W"] = np.einsum('nri,nrj->nij', Je, Jf) * (w * w)[:, None, None]
    return out


def oracle_reduced(p, elim_view, radius, min_diag=1e-6, max_diag=1e32):
    """(S, b, f_index, H, g, d2) with LM damping on the eliminated blocks only and
    constant eliminated blocks decoupled -- what rcc_ba_schur produces."""
    H, g, _ = O.normal_equations(p)
    o_view, o_marker, o_shared, n = p.offsets()
    nv, nm = len(p.views), len(p.markers)
    e_off, n_e = (o_view, nv) if elim_view else (o_marker, nm)
    f_off, n_f = (o_marker, nm) if elim_view else (o_view, nv)
    d2 = np.clip(np.diag(H), min_diag, max_diag) / radius
    Hd = H.copy()
    e_idx = np.arange(e_off, e_off + 6 * n_e)
    Hd[e_idx, e_idx] += d2[e_idx]
    cm = p.const_mask()
    ce = e_idx[cm[e_idx]]
    gd = g.copy()
    Hd[ce, :] = 0
    Hd[:, ce] = 0
    Hd[ce, ce] = 1
    gd[ce] = 0
    f_index = np.concatenate([np.arange(f_off, f_off + 6 * n_f), np.arange(o_shared, n)])
    S, b = O.schur_reduce(Hd, gd, slice(e_off, e_off + 6 * n_e), f_index)
    return S, b, f_index, H, g, d2


def load_golden(name):
    """tests/golden/<name>.npz -> (Scene, dict of expected arrays)."""
    import os
    from robot_camera_calibration_b200.scenes import Scene
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz")))
    s = Scene(model=str(g["model"]), intr=g["intr"], dist=g["dist"], ext=g["ext"], views=g["views"],
              markers=g["markers"], sizes=g["sizes"], view_idx=g["view_idx"], marker_idx=g["marker_idx"],
              cam_idx=g["cam_idx"], pixels=g["pixels"], const_views=g["const_views"],
              const_markers=g["const_markers"], const_intr=g["const_intr"], const_dist=g["const_dist"],
              const_ext=g["const_ext"])
    return s, g


GOLDEN = ("single_small", "rig_small", "single_intpix")
