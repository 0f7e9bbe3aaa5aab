"""ctypes wrapper of the "Ceres-equivalent" CPU restatement (cpu_restatement.cpp).

TEST / BASELINE INFRASTRUCTURE ONLY (see the header of cpu_restatement.cpp):
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by the product package.  Parity status: unpinned
at the Ceres boundary (no Ceres here); pinned to the numpy/OpenCV oracles by
tests/test_cpu_restatement.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libcpu_restatement.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint8)
_lib = None


def build(force=False):
    src = os.path.join(HERE, "cpu_restatement.cpp")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-B", "libcpu_restatement.so"], check=True, capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        lib = C.CDLL(LIB)
        lib.cpu_ba_create.restype = C.c_void_p
        lib.cpu_ba_create.argtypes = [C.c_int] * 4 + [C.c_int64, C.c_int, _ip, _ip, _ip, _dp]
        lib.cpu_ba_destroy.argtypes = [C.c_void_p]
        lib.cpu_ba_set_num_threads.argtypes = [C.c_int]
        lib.cpu_ba_set_params.argtypes = [C.c_void_p] + [_dp] * 6
        lib.cpu_ba_evaluate.argtypes = [C.c_void_p, C.c_int]
        lib.cpu_ba_linearize.argtypes = [C.c_void_p, _dp]
        lib.cpu_ba_get.argtypes = [C.c_void_p] + [_dp] * 15
        lib.cpu_ba_schur.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, _up, _dp, _dp]
        lib.cpu_ba_back_substitute.argtypes = [C.c_void_p, _up, _dp, _dp]
        _lib = lib
    return _lib


def _p(a, t=_dp):
    return None if a is None else a.ctypes.data_as(t)


def set_num_threads(n):
    """Size of the OpenMP team of the restatement (the environment variable is too late once libgomp is loaded)."""
    load().cpu_ba_set_num_threads(int(n))


class CpuBA:
    """CPU restatement of one BA problem (all host threads via OpenMP)."""

    def __init__(self, scene, eliminate="auto"):
        self.lib = load()
        s = scene
        self.model = s.model
        self.rig = int(s.model == "rig")
        self.nv, self.nm, self.nc, self.n = len(s.views), len(s.markers), len(s.intr), len(s.view_idx)
        self.elim_view = (eliminate == "views") or (eliminate == "auto" and self.nv >= self.nm)
        self.vi = np.ascontiguousarray(s.view_idx, np.int32)
        self.mi = np.ascontiguousarray(s.marker_idx, np.int32)
        self.ci = np.ascontiguousarray(s.cam_idx, np.int32)
        self.pix = np.ascontiguousarray(s.pixels, np.float64)
        self.h = self.lib.cpu_ba_create(self.rig, self.nv, self.nm, self.nc, self.n, int(self.elim_view),
                                        _p(self.vi, _ip), _p(self.mi, _ip), _p(self.ci, _ip), _p(self.pix))
        self.sp = 15 if self.rig else 9
        self.ns = self.nc * self.sp
        self.n_e, self.n_f = (self.nv, self.nm) if self.elim_view else (self.nm, self.nv)
        self.views, self.markers = s.views.copy(), s.markers.copy()
        self.intr, self.dist, self.ext, self.sizes = s.intr.copy(), s.dist.copy(), s.ext.copy(), s.sizes.copy()
        self.const_views, self.const_markers = s.const_views.copy(), s.const_markers.copy()
        self.const_intr, self.const_dist, self.const_ext = s.const_intr.copy(), s.const_dist.copy(), s.const_ext.copy()
        self.push_params()

    def push_params(self):
        c = lambda a: np.ascontiguousarray(a, np.float64)
        self._keep = [c(self.views), c(self.markers), c(self.sizes), c(self.intr), c(self.dist), c(self.ext)]
        self.lib.cpu_ba_set_params(self.h, *[_p(a) for a in self._keep])

    def close(self):
        if self.h:
            self.lib.cpu_ba_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def threads(self):
        return int(self.lib.cpu_ba_num_threads())

    def evaluate(self, want_jacobians=True):
        return self.lib.cpu_ba_evaluate(self.h, int(want_jacobians))

    def linearize(self):
        c = C.c_double()
        fail = self.lib.cpu_ba_linearize(self.h, C.byref(c))
        return c.value, fail

    def get(self):
        n, ne, nf, ns = self.n, self.n_e, self.n_f, self.ns
        o = dict(residuals=np.empty((n, 8)), intr=np.empty((n, 8, 4)), dist=np.empty((n, 8, 5)),
                 view=np.empty((n, 8, 6)), marker=np.empty((n, 8, 6)),
                 ext=np.empty((n, 8, 6)) if self.rig else None,
                 Hee=np.empty((ne, 6, 6)), ge=np.empty((ne, 6)), Hes=np.empty((ne, 6, ns)),
                 Hff=np.empty((nf, 6, 6)), gf=np.empty((nf, 6)), Hfs=np.empty((nf, 6, ns)),
                 Hss=np.empty((ns, ns)), gs=np.empty(ns), W=np.empty((n, 6, 6)))
        self.lib.cpu_ba_get(self.h, *[_p(o[k]) for k in ("residuals", "intr", "dist", "view", "marker", "ext", "Hee",
                                                        "ge", "Hes", "Hff", "gf", "Hfs", "Hss", "gs", "W")])
        return o

    def _const_e(self):
        return np.ascontiguousarray(self.const_views if self.elim_view else self.const_markers, np.uint8)

    def _const_reduced(self):
        cf = self.const_markers if self.elim_view else self.const_views
        m = np.repeat(cf, 6)
        sh = []
        for c in range(self.nc):
            sh += [self.const_intr[c]] * 4 + [self.const_dist[c]] * 5
            if self.rig:
                sh += [self.const_ext[c]] * 6
        return np.concatenate([m, np.array(sh, bool)])

    def schur(self, radius, min_diag=1e-6, max_diag=1e32):
        n_red = 6 * self.n_f + self.ns
        S, b = np.empty((n_red, n_red)), np.empty(n_red)
        ce = self._const_e()
        bad = self.lib.cpu_ba_schur(self.h, radius, min_diag, max_diag, _p(ce, _up), _p(S), _p(b))
        return S, b, bad

    def lm_iteration(self, radius, min_diag=1e-6, max_diag=1e32, timings=None):
        """linearize -> Schur -> dense Cholesky (LAPACK via SciPy) -> back-substitution.
        Returns (delta_e, delta_F, cost)."""
        import scipy.linalg as sla
        t0 = time.perf_counter()
        cost, fail = self.linearize()
        t1 = time.perf_counter()
        S, b, bad = self.schur(radius, min_diag, max_diag)
        t2 = time.perf_counter()
        o = self.get() if False else None
        n_red = len(b)
        # F-side damping from the diagonal of J^T J (not of S) + constant mask
        Hff = np.empty((self.n_f, 6, 6)); Hss = np.empty((self.ns, self.ns))
        self.lib.cpu_ba_get(self.h, *([None] * 9 + [_p(Hff)] + [None] * 2 + [_p(Hss)] + [None] * 2))
        hd = np.concatenate([np.einsum('fii->fi', Hff).ravel(), np.diag(Hss)])
        S[np.diag_indices(n_red)] += np.clip(hd, min_diag, max_diag) / radius
        cm = self._const_reduced()
        idx = np.nonzero(cm)[0]
        S[idx, :] = 0; S[:, idx] = 0; S[idx, idx] = 1; b[idx] = 0
        cf = sla.cho_factor(S, lower=True, overwrite_a=True, check_finite=False)
        dF = -sla.cho_solve(cf, b, check_finite=False)
        t3 = time.perf_counter()
        dE = np.empty((self.n_e, 6))
        ce = self._const_e()
        self.lib.cpu_ba_back_substitute(self.h, _p(ce, _up), _p(np.ascontiguousarray(dF)), _p(dE))
        t4 = time.perf_counter()
        if timings is not None:
            timings.update(linearize=t1 - t0, schur=t2 - t1, solve=t3 - t2, backsub=t4 - t3)
        return dE, dF, cost
