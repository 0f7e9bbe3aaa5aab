"""CPU: lane-level restatements of the index arithmetic of the block-sparse Schur SYRK kernels (csrc/schur.cu) against
a direct sum of Y_e^T Y_e -- the presence-mask / popcount addressing of schur_syrk_reg_kernel (v3) and the MMA fragment
layout, column-group decomposition and slice addressing of schur_syrk_mma_kernel (v4).  The kernels themselves are
checked on the GPU (tests/test_gpu_parity.py, test_gpu_edge_cases.py); these keep the host-built tables (tile_ptr,
tile_mask: problem.cu) and the per-lane formulas honest without one."""
import numpy as np


def _structure(n_e, n_f, dens, seed):
    rng = np.random.default_rng(seed)
    rows = [np.sort(rng.choice(n_f, size=max(1, rng.binomial(n_f, dens)), replace=False)) for _ in range(n_e)]
    pair_e = np.concatenate([[e] * len(r) for e, r in enumerate(rows)])
    pair_f = np.concatenate(rows)
    row_ptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])])
    Y = rng.standard_normal((len(pair_e), 36))           # Y[p][6 col + k] = Y_p[k][col]  (column-major 6 x 6 records)
    nt = (n_f + 31) // 32
    tile_ptr = np.zeros((n_e, nt + 1), int)
    tile_mask = np.zeros((n_e, nt), np.uint64)
    for e in range(n_e):                                  # problem.cu: tile_ptr / tile_mask
        p, end = row_ptr[e], row_ptr[e + 1]
        for J in range(nt + 1):
            while p < end and pair_f[p] < 32 * J:
                p += 1
            tile_ptr[e, J] = end if J == nt else p
        for p in range(row_ptr[e], row_ptr[e + 1]):
            tile_mask[e, pair_f[p] // 32] |= np.uint64(1) << np.uint64(pair_f[p] % 32)
    S = np.zeros((6 * n_f, 6 * n_f))
    for e in range(n_e):
        ps = range(row_ptr[e], row_ptr[e + 1])
        for p in ps:
            A = Y[p].reshape(6, 6).T
            for q in ps:
                S[6 * pair_f[p]:6 * pair_f[p] + 6, 6 * pair_f[q]:6 * pair_f[q] + 6] += A.T @ Y[q].reshape(6, 6).T
    return pair_e, pair_f, Y, nt, tile_ptr, tile_mask, S


def _check_upper(got, S, n_f):
    for f in range(n_f):
        for fp in range(f, n_f):
            np.testing.assert_allclose(got[6 * f:6 * f + 6, 6 * fp:6 * fp + 6], S[6 * f:6 * f + 6, 6 * fp:6 * fp + 6],
                                       rtol=1e-12, atol=1e-12)


def _window(j):                                           # schur.cu: sr_window
    b0, b1 = (32 * j) // 6, (32 * j + 31) // 6
    hi = 0xffffffff if b1 >= 31 else (1 << (b1 + 1)) - 1
    return hi & ~((1 << b0) - 1) & 0xffffffff


def test_register_accumulator_kernel_addressing():
    n_e, n_f = 24, 70
    pair_e, pair_f, Y, nt, tile_ptr, tile_mask, S = _structure(n_e, n_f, 0.3, 1)
    assert [hex(_window(j)) for j in range(6)] == ['0x3f', '0x7e0', '0xfc00', '0x3f0000', '0x7e00000', '0xfc000000']
    got = np.full_like(S, np.nan)
    for f in range(n_f):
        col_pairs = np.nonzero(pair_f == f)[0]            # ascending e
        for js in range(f // 32, nt):
            keep = (~((1 << (f & 31)) - 1)) & 0xffffffff if js == f // 32 else 0xffffffff
            acc = np.zeros((32, 6, 6))
            for p in col_pairs:
                e = pair_e[p]
                mfull = int(tile_mask[e, js])
                m, base = mfull & keep, tile_ptr[e, js]
                if m == 0:
                    continue
                for lane in range(32):
                    for j in range(6):
                        if not (m & _window(j)):
                            continue
                        c = lane + 32 * j
                        blk = c // 6
                        colo = (c - blk * 6) * 6
                        y = np.zeros(6)
                        if (m >> blk) & 1:
                            idx = base + bin(mfull & ((1 << blk) - 1)).count("1")
                            assert pair_e[idx] == e and pair_f[idx] == 32 * js + blk
                            y = Y[idx, colo:colo + 6]
                        acc[lane, j] += Y[p].reshape(6, 6) @ y
            for lane in range(32):
                for j in range(6):
                    c = lane + 32 * j
                    fp = 32 * js + c // 6
                    if f <= fp < n_f:
                        got[6 * f:6 * f + 6, 6 * 32 * js + c] = acc[lane, j]
    _check_upper(got, S, n_f)


def _mma(a, b):
    """mma.m8n8k4.f64 on lane values: lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4)], D[l/4][2(l%4)+1]."""
    A, B = np.zeros((8, 4)), np.zeros((4, 8))
    for l in range(32):
        A[l // 4][l % 4] = a[l]
        B[l % 4][l // 4] = b[l]
    D = A @ B
    return np.array([[D[l // 4][2 * (l % 4)], D[l // 4][2 * (l % 4) + 1]] for l in range(32)])


def test_tensor_core_product_kernel_fragments_and_addressing():
    n_e, n_f = 14, 70
    pair_e, pair_f, Y, nt, tile_ptr, _, S = _structure(n_e, n_f, 0.45, 3)
    Yflat = Y.ravel()
    got = np.full_like(S, np.nan)
    max_groups = 0
    for f in range(n_f):
        col_pairs = np.nonzero(pair_f == f)[0]
        sub_of_f = f >> 5
        for js in range(sub_of_f, nt):
            subbase = 32 * js
            acc = np.zeros(32 * 36)                       # [block][row][column]
            for p in col_pairs:
                e = pair_e[p]
                lo = p if js == sub_of_f else tile_ptr[e, js]
                hi = tile_ptr[e, js + 1]
                n_p = max(hi - lo, 0)
                ncol = 6 * n_p
                ngrp = (ncol + 7) >> 3
                max_groups = max(max_groups, ngrp)
                pf_lane = [(pair_f[lo + l] - subbase) if l < n_p else 0 for l in range(32)]
                yb = Y[p]
                a0 = [yb[(l >> 2) * 6 + (l & 3)] if (l >> 2) < 6 else 0.0 for l in range(32)]
                a1 = [yb[(l >> 2) * 6 + 4 + (l & 3)] if ((l >> 2) < 6 and (l & 3) < 2) else 0.0 for l in range(32)]
                for cg in range(ngrp):
                    b0, b1 = np.zeros(32), np.zeros(32)
                    for l in range(32):
                        g, t = l >> 2, l & 3
                        col = 8 * cg + g
                        if col < ncol:
                            b0[l] = Yflat[lo * 36 + col * 6 + t]
                            if t < 2:
                                b1[l] = Yflat[lo * 36 + col * 6 + 4 + t]
                    d = _mma(a0, b0) + _mma(a1, b1)
                    for l in range(32):
                        g, t = l >> 2, l & 3
                        col = 8 * cg + 2 * t
                        jj = col // 6
                        c = col - 6 * jj
                        if g < 6 and col < ncol:
                            o = pf_lane[jj & 31] * 36 + g * 6 + c
                            acc[o] += d[l][0]
                            acc[o + 1] += d[l][1]
            for r in range(6):
                for cl in range(min(32, n_f - subbase) * 6):
                    fl = cl // 6
                    if subbase + fl >= f:
                        got[6 * f + r, 6 * subbase + cl] = acc[fl * 36 + r * 6 + (cl - fl * 6)]
    assert max_groups > 6                                 # rows with more than one chunk of column groups
    _check_upper(got, S, n_f)
