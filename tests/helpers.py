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


# ---------------------------------------------------------------------------------------------------
# BASELINE-size oracle: the same blocks as oracle_blocks() / oracle_reduced() without ever forming the
# dense Jacobian or the dense normal matrix (vectorised complex step per block + einsum + np.add.at)
# ---------------------------------------------------------------------------------------------------
def _segment_sum(idx, n, X):
    """out[i] = sum of X[k] over idx[k] == i  (np.add.at semantics, at sparse-matmul speed)."""
    import scipy.sparse as sp
    N = len(idx)
    M = sp.csr_matrix((np.ones(N), (idx, np.arange(N))), shape=(n, N))
    return np.asarray(M @ X.reshape(N, -1)).reshape((n,) + X.shape[1:])


def oracle_blocks_sparse(p, elim_view, slab=40000, threads=None):
    """Normal-equation blocks in the GPU library's layout from ba_oracle's per-block Jacobians.
    Same keys as oracle_blocks() minus the dense H / g.  Works in slabs of observation blocks (one host thread
    each: numpy releases the GIL in the heavy loops) so that cfg2 at full size (0.46 M blocks) takes seconds."""
    import copy
    import os
    from concurrent.futures import ThreadPoolExecutor
    nv, nm, nc, sp = len(p.views), len(p.markers), p.n_cam, p.shared_per_cam
    n_e, n_f = (nv, nm) if elim_view else (nm, nv)
    ns = nc * sp
    N = len(p.view_idx)
    out = {"W": np.empty((N, 6, 6)), "residuals": np.empty((N, 8)), "n_e": n_e, "n_f": n_f, "n_shared": ns}
    Jall = {k: np.empty((N, 8, w)) for k, w in O.local_param_names(p.model)}
    JsJe = np.empty((N, 6, sp))      # per block  J_e^T J_shared, J_f^T J_shared, J_shared^T J_shared, gradients
    JsJf = np.empty((N, 6, sp))
    JsJs = np.empty((N, sp, sp))
    Hee_b, Hff_b = np.empty((N, 6, 6)), np.empty((N, 6, 6))
    ge_b, gf_b, gs_b = np.empty((N, 6)), np.empty((N, 6)), np.empty((N, sp))
    rho = np.empty(N)

    def work(a):
        b = min(N, a + slab)
        q = copy.copy(p)
        q.view_idx, q.marker_idx, q.cam_idx, q.pixels = p.view_idx[a:b], p.marker_idx[a:b], p.cam_idx[a:b], p.pixels[a:b]
        r = O.residuals(q)
        Jb = O.jacobian_blocks_cs(q)
        out["residuals"][a:b] = r
        for k in Jb:
            Jall[k][a:b] = Jb[k]
        w, rho[a:b] = O.robust(q, r)
        rw = r * w[:, None]
        Jw = {k: v * w[:, None, None] for k, v in Jb.items()}
        Je, Jf = (Jw["view"], Jw["marker"]) if elim_view else (Jw["marker"], Jw["view"])
        Js = np.concatenate([Jw["intr"], Jw["dist"]] + ([Jw["ext"]] if p.model == "rig" else []), axis=2)
        Hee_b[a:b] = np.einsum('nri,nrj->nij', Je, Je)
        Hff_b[a:b] = np.einsum('nri,nrj->nij', Jf, Jf)
        out["W"][a:b] = np.einsum('nri,nrj->nij', Je, Jf)
        JsJe[a:b] = np.einsum('nri,nrj->nij', Je, Js)
        JsJf[a:b] = np.einsum('nri,nrj->nij', Jf, Js)
        JsJs[a:b] = np.einsum('nri,nrj->nij', Js, Js)
        ge_b[a:b] = np.einsum('nri,nr->ni', Je, rw)
        gf_b[a:b] = np.einsum('nri,nr->ni', Jf, rw)
        gs_b[a:b] = np.einsum('nri,nr->ni', Js, rw)

    if threads is None:
        threads = min(len(os.sched_getaffinity(0)), 16)
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex:
        list(ex.map(work, range(0, N, slab)))
    ei, fi = (p.view_idx, p.marker_idx) if elim_view else (p.marker_idx, p.view_idx)
    ci = p.cam_idx
    out["Hee"], out["ge"] = _segment_sum(ei, n_e, Hee_b), _segment_sum(ei, n_e, ge_b)
    out["Hff"], out["gf"] = _segment_sum(fi, n_f, Hff_b), _segment_sum(fi, n_f, gf_b)
    # shared columns: camera c owns columns c*sp .. c*sp+sp-1
    Hes = _segment_sum(ei * nc + ci, n_e * nc, JsJe).reshape(n_e, nc, 6, sp)
    Hfs = _segment_sum(fi * nc + ci, n_f * nc, JsJf).reshape(n_f, nc, 6, sp)
    out["Hes"] = np.ascontiguousarray(Hes.transpose(0, 2, 1, 3)).reshape(n_e, 6, ns)
    out["Hfs"] = np.ascontiguousarray(Hfs.transpose(0, 2, 1, 3)).reshape(n_f, 6, ns)
    Hss_c, gs_c = _segment_sum(ci, nc, JsJs), _segment_sum(ci, nc, gs_b)
    out["Hss"] = np.zeros((ns, ns))
    for c in range(nc):
        out["Hss"][c * sp:(c + 1) * sp, c * sp:(c + 1) * sp] = Hss_c[c]
    out["gs"] = gs_c.reshape(ns)
    out["cost"] = 0.5 * float(rho.sum())
    out["jacobians"] = Jall
    return out


class SparseSchurOracle:
    """Blocks of the reduced system  S = H_FF - H_FE (H_EE + D)^-1 H_EF,  b = g_F - H_FE (H_EE + D)^-1 g_E
    (damping on the eliminated blocks only, constant eliminated blocks decoupled: what rcc_ba_schur produces),
    computed for chosen kept blocks from oracle_blocks_sparse() -- never the whole matrix."""

    def __init__(self, p, ob, elim_view, radius, min_diag=1e-6, max_diag=1e32):
        import scipy.sparse as sp
        self.ob, self.n_f, self.ns = ob, ob["n_f"], ob["n_shared"]
        n_e = ob["n_e"]
        ei, fi = (p.view_idx, p.marker_idx) if elim_view else (p.marker_idx, p.view_idx)
        ce = p.const_views if elim_view else p.const_markers
        d2 = np.clip(np.einsum('eii->ei', ob["Hee"]), min_diag, max_diag) / radius
        A = ob["Hee"] + np.einsum('ei,ij->eij', d2, np.eye(6))
        self.Ainv = np.linalg.inv(A)
        self.Ainv[np.asarray(ce, bool)] = 0.0           # constant eliminated block: no coupling
        key = ei.astype(np.int64) * self.n_f + fi
        uniq, inv = np.unique(key, return_inverse=True)
        self.Wp = np.zeros((len(uniq), 6, 6))
        np.add.at(self.Wp, inv, ob["W"])                 # several cameras may see the same (e, f)
        pe, pf = uniq // self.n_f, uniq % self.n_f
        self.by_f = sp.csr_matrix((np.arange(1, len(uniq) + 1), (pf, pe)), shape=(self.n_f, n_e))
        self.border_e = np.concatenate([ob["Hes"], ob["ge"][:, :, None]], axis=2)      # (n_e, 6, ns+1)

    def _row(self, f):
        r = self.by_f.getrow(f)
        return r.indices, r.data - 1

    def block(self, f, g):
        """S[6f:6f+6, 6g:6g+6]"""
        e1, p1 = self._row(f)
        e2, p2 = self._row(g)
        _, i1, i2 = np.intersect1d(e1, e2, return_indices=True)
        e = e1[i1]
        acc = np.einsum('eki,ekl,elj->ij', self.Wp[p1[i1]], self.Ainv[e], self.Wp[p2[i2]])
        base = self.ob["Hff"][f] if f == g else 0.0
        return base - acc

    def border(self, f):
        """[S[6f:6f+6, shared] | b[6f:6f+6]]   (6, ns+1)"""
        e, pid = self._row(f)
        acc = np.einsum('eki,ekl,elj->ij', self.Wp[pid], self.Ainv[e], self.border_e[e])
        base = np.concatenate([self.ob["Hfs"][f], self.ob["gf"][f][:, None]], axis=1)
        return base - acc

    def corner(self):
        """[S[shared, shared] | b[shared]]   (ns, ns+1)"""
        acc = np.einsum('eki,ekl,elj->ij', self.border_e[:, :, :self.ns], self.Ainv, self.border_e)
        base = np.concatenate([self.ob["Hss"], self.ob["gs"][:, None]], axis=1)
        return base - acc
