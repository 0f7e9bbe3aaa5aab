"""GPU: the hand-written dense Cholesky of the reduced solve (csrc/dense.cu) against LAPACK semantics
(torch.linalg.cholesky = cuSOLVER / MAGMA potrf) and against cusolverDnDpotrf through the same C entry point."""
import ctypes as C

import numpy as np
import pytest

from robot_camera_calibration_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _spd(n, seed, cond_boost=0.0):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(n, n // 3 + 8, dtype=torch.float64, device="cuda", generator=g)
    S = A @ A.T + (1e-3 + cond_boost) * n * torch.eye(n, dtype=torch.float64, device="cuda")
    return S


def _potrf(buf, n, ld, extra, cusolver=False):
    info, ms = C.c_int32(-1), C.c_double()
    rc = L.load().rcc_dense_potrf(0, C.c_void_p(buf.data_ptr()), n, ld, extra, int(cusolver), C.byref(info), C.byref(ms))
    assert rc == L.RCC_OK
    return info.value, ms.value


@pytest.mark.parametrize("n", [5, 31, 32, 33, 127, 128, 129, 255, 300, 515, 1000, 3009])
def test_own_potrf_matches_lapack(n):
    """Lower-triangular factor and the forward-substituted bordered right-hand side, for sizes around every
    blocking boundary (32-column sub-blocks, 128-column panels, 64-row strips, 128 x 64 update tiles)."""
    import torch
    S = _spd(n, n)
    b = torch.randn(n, dtype=torch.float64, device="cuda")
    ld = n + 1 + ((n + 1) & 1) + 2                         # even, > n + 1
    buf = torch.full((n + 3, ld), float("nan"), dtype=torch.float64, device="cuda")   # NaN: nothing outside may be read
    buf[:n, :n] = torch.triu(S) + torch.tril(torch.full_like(S, float("nan")), -1)   # only the stored triangle is valid
    buf[:n, n] = b
    buf[:n, n + 1:] = 0.0                                  # padding columns of the border are finite in the product
    info, _ = _potrf(buf, n, ld, 1)
    assert info == 0
    Lref = torch.linalg.cholesky(S)
    Lgot = torch.triu(buf[:n, :n]).T
    assert torch.isfinite(Lgot).all()
    err = (torch.linalg.norm(Lgot - Lref) / torch.linalg.norm(Lref)).item()
    assert err < 1e-12, err
    y = torch.linalg.solve_triangular(Lref, b[:, None], upper=False)[:, 0]
    assert (torch.linalg.norm(buf[:n, n] - y) / torch.linalg.norm(y)).item() < 1e-11
    # backward error of the factorisation itself
    assert (torch.linalg.norm(Lgot @ Lgot.T - S) / torch.linalg.norm(S)).item() < 1e-14


def test_own_potrf_reports_the_first_bad_pivot():
    import torch
    n = 300
    S = _spd(n, 7)
    S[200, 200] = -1.0
    buf = torch.zeros((n + 3, n + 4), dtype=torch.float64, device="cuda")
    buf[:n, :n] = S
    info, _ = _potrf(buf, n, n + 4, 1)
    ref = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    ref[:] = S
    info_ref, _ = _potrf(ref, n, n, 0, cusolver=True)
    assert info == info_ref == 201


def test_own_potrf_agrees_with_cusolver_on_a_reduced_system_sized_matrix():
    import torch
    n = 6030
    S = _spd(n, 3)
    a = S.clone()
    b = torch.zeros((n + 3, n + 2), dtype=torch.float64, device="cuda")
    b[:n, :n] = S
    i1, ms_cus = _potrf(a, n, n, 0, cusolver=True)
    i2, ms_own = _potrf(b, n, n + 2, 1)
    assert i1 == 0 and i2 == 0
    La, Lb = torch.triu(a).T, torch.triu(b[:n, :n]).T
    assert (torch.linalg.norm(La - Lb) / torch.linalg.norm(La)).item() < 1e-12
    print(f"n={n}: cusolver {ms_cus:.2f} ms, own {ms_own:.2f} ms")


@pytest.mark.parametrize("n", [5, 31, 32, 33, 127, 128, 129, 255, 300, 515, 1000, 3009, 6030])
def test_own_back_substitution_matches_lapack(n):
    """L^T x = r with the hand-written kernel (one CTA per 128-row block chained through release / acquire flags)
    against torch's triangular solve and against cublasDtrsv through the same entry point."""
    import torch
    S = _spd(n, 100 + n)
    ld = n + 2 + (n & 1)
    buf = torch.full((n + 1, ld), float("nan"), dtype=torch.float64, device="cuda")
    buf[:n, :n] = torch.triu(S) + torch.tril(torch.full_like(S, float("nan")), -1)
    buf[:n, n:] = 0.0
    info, _ = _potrf(buf, n, ld, 1)
    assert info == 0
    r = torch.randn(n, dtype=torch.float64, device="cuda")
    Lf = torch.triu(buf[:n, :n]).T.contiguous()
    want = torch.linalg.solve_triangular(Lf.T, r[:, None], upper=True)[:, 0]
    outs = {}
    for name, cub in (("own", 0), ("cublas", 1)):
        x = r.clone()
        ms = C.c_double()
        rc = L.load().rcc_dense_trsv(0, C.c_void_p(buf.data_ptr()), n, ld, C.c_void_p(x.data_ptr()), cub, C.byref(ms))
        assert rc == L.RCC_OK
        outs[name] = (x, ms.value)
        assert torch.isfinite(x).all()
        assert (torch.linalg.norm(x - want) / torch.linalg.norm(want)).item() < 1e-11, name
    if n >= 3009:
        print(f"n={n}: own trsv {outs['own'][1]:.3f} ms, cublasDtrsv {outs['cublas'][1]:.3f} ms")
