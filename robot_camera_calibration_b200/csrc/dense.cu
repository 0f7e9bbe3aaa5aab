// K5 -- dense Cholesky of the reduced camera system (the "single dense Cholesky solve" of the path), hand-written
// for sm_100a: right-looking blocked factorisation with 128-column panels whose trailing update runs on the FP64
// tensor cores (mma.sync.m8n8k4.f64, SASS DMMA.8x8x4), and block-column-cyclic ownership over the ranks of a
// multi-GPU solve (problem.cu drives the panels, their broadcasts and the look-ahead).
//
// Replaces the dense factorisation that would run inside Ceres (DENSE_SCHUR ->
// Eigen / LAPACK potrf) for the reference's missing optimiser stage; round 1 called cusolverDnDpotrf here, which is
// kept only as the comparator of tests/test_gpu_dense.py.
//
// Storage: the reduced buffer is row-major and holds the UPPER triangle (+ the right-hand side in column n), which
// is the LOWER triangle in column-major order with leading dimension ld: element (i, j), i >= j, sits at A[j*ld+i].
// Rows n .. n_rows-1 (the bordered right-hand side, see problem.cu do_step) ride along: they are rows of every
// panel, never columns.  Three kernels per panel [k0, k0 + kb):
//   chol_diag_kernel    one CTA: the 128 x 128 diagonal block, left-looking over 32-column sub-blocks; a warp
//                       factors each 32 x 32 diagonal sub-block in registers (lane = row, shuffles) and inverts it;
//                       DMMA for everything off the diagonal.  Leaves L_kk in A, a zero/identity-padded copy and the
//                       four 32 x 32 inverses in a workspace.
//   chol_panel_kernel   64-row strips below the block: X <- X L_kk^-T by 32-column sub-steps
//                       X_c <- (X_c - sum_{p<c} X_p L_cp^T) inv(L_cc)^T, all DMMA, the strip resident in shared memory.
//   chol_update_kernel  trailing update C -= P P^T restricted to the block columns this rank owns: 128 x 64 tiles,
//                       8 warps x (32 x 32) accumulators in registers, cp.async 4-stage pipeline over the 128-deep
//                       panel, 2 CTAs per SM (16 warps: the DMMA pipe needs ~4 warps per sub-partition to stay
//                       full) so that one CTA's tile load / write-back overlaps the other's MMAs.
#include "common.cuh"
#include "dense.h"

#include <string>

namespace rcc {

namespace {

constexpr int NB = CHOL_NB;   // panel width
constexpr int SB = 32;        // sub-block of the panel kernels
constexpr int LST = NB + 4;   // shared-memory column stride of a 128-row tile (== 4 mod 16: conflict-free fragments)
constexpr int XST = 64 + 4;   // same for a 64-row tile
constexpr int IST = SB + 4;   // same for a 32-row tile

// D(8x8) += A(8x4) B(4x8), FP64.  Lane l holds A[l/4][l%4], B[l%4][l/4], D[l/4][2(l%4)], D[l/4][2(l%4)+1].
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}
// (mma.m16n8k8.f64 is no shortcut on sm_100a: ptxas lowers it to four DMMA.8x8x4.)
// 16-byte global -> shared copy; bytes == 0 writes zeros without touching src
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------------------
// acc[MT][NT] (8x8 tiles) += A(rows r0.., k) * B(cols c0.., k)^T over k in [0, K), K a multiple of 4.
//   a_sm: A(r, k) at a_sm[k * a_st + r]      b_sm: B(c, k) at b_sm[k * b_st + c]
// ---------------------------------------------------------------------------------------------------------
template <int MT, int NT>
__device__ __forceinline__ void warp_mma(double (&acc)[MT][NT][2], const double* a_sm, int a_st, const double* b_sm,
                                         int b_st, int K, int lane) {
  const int lk = lane & 3, lr = lane >> 2;
  for (int k = 0; k < K; k += 4) {
    double af[MT], bf[NT];
#pragma unroll
    for (int i = 0; i < MT; ++i) af[i] = a_sm[(k + lk) * a_st + 8 * i + lr];
#pragma unroll
    for (int j = 0; j < NT; ++j) bf[j] = b_sm[(k + lk) * b_st + 8 * j + lr];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) dmma(acc[i][j], af[i], bf[j]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// 32 x 32 Cholesky + inverse of the factor by one warp.  blk: column-major, stride st, element (i, j) at blk[j*st+i]
// (lower triangle read and overwritten).  inv: inv(L)(n, k) at inv[k * IST + n] (lower; the rest is zeroed).
// Returns 0 or 1 + the first column whose pivot is not positive.
// ---------------------------------------------------------------------------------------------------------
__device__ int warp_potrf32(double* blk, int st, double* inv, int lane) {
  double a[SB];
#pragma unroll
  for (int j = 0; j < SB; ++j) a[j] = blk[j * st + lane];     // lane = row
  double rdiag = 1.0;
  int bad = 0;
#pragma unroll
  for (int j = 0; j < SB; ++j) {
    double d = __shfl_sync(0xffffffffu, a[j], j);
    if (!(d > 0.0)) {              // warp-uniform
      if (!bad) bad = j + 1;
      d = 1.0;
    }
    const double r = rsqrt(d);
    const double lj = a[j] * r;    // L(i, j) for lanes i >= j (lane j: sqrt(d) = d / sqrt(d))
    a[j] = lj;
    if (lane == j) rdiag = r;
#pragma unroll
    for (int k = j + 1; k < SB; ++k) {
      const double lk = __shfl_sync(0xffffffffu, lj, k);
      a[k] = fma(-lj, lk, a[k]);
    }
  }
#pragma unroll
  for (int j = 0; j < SB; ++j)
    if (lane >= j) blk[j * st + lane] = a[j];
  __syncwarp();
  // inverse, lane = column c of M = L^-1:  M(i, c) = (delta_ic - sum_{p<i} L(i, p) M(p, c)) / L(i, i); M(p, c) = 0 for p < c
  double m[SB];
#pragma unroll
  for (int i = 0; i < SB; ++i) {
    double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;    // four chains: FMA latency, not 4x the work
#pragma unroll
    for (int p = 0; p < i; ++p) {                                        // blk[..]: uniform address = broadcast
      const double l = blk[p * st + i];
      if ((p & 3) == 0) s0 = fma(-l, m[p], s0);
      else if ((p & 3) == 1) s1 = fma(-l, m[p], s1);
      else if ((p & 3) == 2) s2 = fma(-l, m[p], s2);
      else s3 = fma(-l, m[p], s3);
    }
    m[i] = ((s0 + s1) + (s2 + s3)) * __shfl_sync(0xffffffffu, rdiag, i);
  }
#pragma unroll
  for (int i = 0; i < SB; ++i) inv[lane * IST + i] = m[i];               // inv(n = i, k = lane)
  __syncwarp();
  return bad;
}

// ---------------------------------------------------------------------------------------------------------
// diagonal block
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) chol_diag_kernel(double* __restrict__ A, int ld, int k0, int kb,
                                                          double* __restrict__ Lw, double* __restrict__ inv32,
                                                          int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* Ls = sm;                       // [NB][LST] column-major
  double* Is = sm + NB * LST;            // [SB][IST] inverse of the current diagonal sub-block
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // load the block (asynchronous 16-byte copies, all in flight at once: a plain load loop would pay the memory
  // latency 128 times); the strict upper triangle comes along and is never used; rows/columns >= kb are padded
  // with the identity
  for (int idx = tid; idx < NB * (NB / 2); idx += 128) {
    const int j = idx >> 6, i2 = (idx & 63) * 2;
    const bool ok = j < kb && i2 + 1 >= j && i2 < kb;
    cp_async16(Ls + j * LST + i2, ok ? A + (size_t)(k0 + j) * ld + k0 + i2 : A, ok ? 16 : 0);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  if (kb < NB) {
    if (tid >= kb) Ls[tid * LST + tid] = 1.0;
    if ((kb & 1) && tid < kb) Ls[tid * LST + kb] = 0.0;      // the pair (kb-1, kb) brought one row too many
    __syncthreads();
  }
  for (int c = 0; c < NB / SB; ++c) {
    const int c0 = c * SB;
    // (1) left-looking: rows >= c0 of block column c  -=  L(rows, 0:c0) L(c0:c0+32, 0:c0)^T ; warp w takes the
    //     8-row tiles w, w + 4, ...
    if (c > 0) {
      const int mt = (NB - c0) / 8;
      for (int t = warp; t < mt; t += 4) {
        double acc[1][4][2] = {};
        warp_mma<1, 4>(acc, Ls + c0 + 8 * t, LST, Ls + c0, LST, c0, lane);
        const int r = c0 + 8 * t + (lane >> 2);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int col = c0 + 8 * j + 2 * (lane & 3) + q;
            Ls[col * LST + r] -= acc[0][j][q];
          }
      }
    }
    __syncthreads();
    // (2) the 32 x 32 diagonal sub-block: factor + invert in registers
    if (warp == 0) {
      const int bad = warp_potrf32(Ls + c0 * LST + c0, LST, Is, lane);
      if (bad && lane == 0) atomicCAS(info, 0, k0 + c0 + bad);
    }
    __syncthreads();
    for (int idx = tid; idx < SB * SB; idx += 128) {
      const int k = idx >> 5, n = idx & 31;
      inv32[c * SB * SB + idx] = Is[k * IST + n];            // inv(n, k) at [k * 32 + n]
    }
    // (3) rows below the sub-block:  X <- X inv(L_cc)^T   (each warp owns whole 8-row tiles: no cross-warp hazard)
    const int mt = (NB - c0 - SB) / 8;
    for (int t = warp; t < mt; t += 4) {
      double acc[1][4][2] = {};
      warp_mma<1, 4>(acc, Ls + c0 * LST + c0 + SB + 8 * t, LST, Is, IST, SB, lane);
      __syncwarp();
      const int r = c0 + SB + 8 * t + (lane >> 2);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) Ls[(c0 + 8 * j + 2 * (lane & 3) + q) * LST + r] = acc[0][j][q];
    }
    __syncthreads();
  }
  // write back: the real part into A, the padded block into the workspace (upper triangle zero)
  for (int idx = tid; idx < NB * NB; idx += 128) {
    const int j = idx >> 7, i = idx & (NB - 1);
    const double v = (i >= j) ? Ls[j * LST + i] : 0.0;
    Lw[idx] = v;
    if (i >= j && i < kb) A[(size_t)(k0 + j) * ld + k0 + i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// panel rows below the diagonal block
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) chol_panel_kernel(double* __restrict__ A, int ld, int n_rows, int k0, int kb,
                                                           int row0, const double* __restrict__ Lw,
                                                           const double* __restrict__ inv32) {
  extern __shared__ __align__(16) double sm[];
  double* Xs = sm;                               // [NB cols][XST]: strip, X(r, c) at Xs[c * XST + r]
  double* Lb = Xs + NB * XST;                    // [6][SB][IST]: L_cp (c > p): L(32c+n, 32p+k) at [blk][k * IST + n]
  double* Ib = Lb + 6 * SB * IST;                // [4][SB][IST]: inv(L_cc)(n, k) at [c][k * IST + n]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // strips start on an even row (16-byte copies); a row below row0 that this pulls in belongs to the diagonal
  // block: it is computed on and never stored
  const int i0 = (row0 & ~1) + 64 * blockIdx.x;
  for (int idx = tid; idx < NB * 32; idx += 128) {           // 64 rows = 32 double2 per column
    const int c = idx >> 5, r2 = (idx & 31) * 2;
    const bool ok = c < kb && i0 + r2 < n_rows;
    cp_async16(Xs + c * XST + r2, ok ? A + (size_t)(k0 + c) * ld + i0 + r2 : A, ok ? 16 : 0);
  }
  cp_async_commit();
  for (int idx = tid; idx < 6 * SB * (SB / 2); idx += 128) {
    const int blk = idx >> 9, k = (idx >> 4) & 31, n = (idx & 15) * 2;
    const int c = blk < 1 ? 1 : (blk < 3 ? 2 : 3), p = blk - (c * (c - 1)) / 2;
    cp_async16(Lb + blk * SB * IST + k * IST + n, Lw + (SB * p + k) * NB + SB * c + n, 16);
  }
  for (int idx = tid; idx < 4 * SB * (SB / 2); idx += 128) {
    const int c = idx >> 9, k = (idx >> 4) & 31, n = (idx & 15) * 2;
    cp_async16(Ib + c * SB * IST + k * IST + n, inv32 + c * SB * SB + k * SB + n, 16);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();
  // warp w owns rows 16w .. 16w+15 of the strip through all four sub-steps
  const int r0 = 16 * warp;
  const int lr = lane >> 2, lc = 2 * (lane & 3);
  for (int c = 0; c < 4; ++c) {
    if (c > 0) {
      double acc[2][4][2] = {};
      for (int p = 0; p < c; ++p)
        warp_mma<2, 4>(acc, Xs + (SB * p) * XST + r0, XST, Lb + ((c * (c - 1)) / 2 + p) * SB * IST, IST, SB, lane);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int q = 0; q < 2; ++q) Xs[(SB * c + 8 * j + lc + q) * XST + r0 + 8 * i + lr] -= acc[i][j][q];
      __syncwarp();
    }
    double acc[2][4][2] = {};
    warp_mma<2, 4>(acc, Xs + (SB * c) * XST + r0, XST, Ib + c * SB * IST, IST, SB, lane);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 2; ++q) Xs[(SB * c + 8 * j + lc + q) * XST + r0 + 8 * i + lr] = acc[i][j][q];
    __syncwarp();
  }
  __syncthreads();
  for (int idx = tid; idx < NB * 32; idx += 128) {
    const int c = idx >> 5, r2 = (idx & 31) * 2;
    if (c >= kb) continue;
    double* dst = A + (size_t)(k0 + c) * ld + i0 + r2;
    const double2 v = *reinterpret_cast<const double2*>(Xs + c * XST + r2);
    if (i0 + r2 >= row0 && i0 + r2 + 1 < n_rows) *reinterpret_cast<double2*>(dst) = v;
    else {
      if (i0 + r2 >= row0 && i0 + r2 < n_rows) dst[0] = v.x;
      if (i0 + r2 + 1 >= row0 && i0 + r2 + 1 < n_rows) dst[1] = v.y;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// trailing update
// ---------------------------------------------------------------------------------------------------------
constexpr int UP_BM = 128, UP_BN = 64, UP_BK = 16, UP_STAGES = 4;
constexpr int UP_AST = UP_BM + 4, UP_BST = UP_BN + 4;
constexpr int UP_STAGE_DOUBLES = UP_BK * (UP_AST + UP_BST);

struct UpdateArgs {
  double* A;
  int ld, n_rows, n_cols;   // rows 0 .. n_rows-1 exist; columns >= n_cols are never updated
  int k0, kb;               // panel
  int first_blk, blk_stride;  // block columns first_blk, first_blk + blk_stride, ... (128 wide)
  int n_tiles_m;            // 128-row tiles of the whole matrix: ceil(n_rows / 128)
};

// CTA id -> (block column b, 64-column half, 128-row tile) over the lower-triangular part only: block column b
// (global block index t_b = first_blk + b * blk_stride) has n_tiles_m - t_b row tiles, each split into two halves.
// P(b) = CTAs before block column b = 2 [b A - S b (b - 1) / 2],  A = n_tiles_m - first_blk,  S = blk_stride.
__device__ __forceinline__ void update_tile_of(const UpdateArgs& a, int id, int& j0, int& i0) {
  const double A = (double)(a.n_tiles_m - a.first_blk), S = (double)a.blk_stride;
  const double disc = (2.0 * A + S) * (2.0 * A + S) - 4.0 * S * (double)id;
  int b = (int)(((2.0 * A + S) - sqrt(fmax(disc, 0.0))) / (2.0 * S));
  auto P = [&](int bb) { return 2 * (bb * (a.n_tiles_m - a.first_blk) - a.blk_stride * ((bb * (bb - 1)) / 2)); };
  while (b > 0 && P(b) > id) --b;
  while (P(b + 1) <= id) ++b;
  const int tb = a.first_blk + b * a.blk_stride;
  const int cnt = a.n_tiles_m - tb;
  const int rem = id - P(b);
  const int half = rem / cnt, it = rem - half * cnt;
  j0 = tb * NB + half * UP_BN;
  i0 = (tb + it) * UP_BM;
}

// 8 warps: warp (wm, wn) owns the 32 x 32 sub-tile at rows 32 wm, columns 32 wn of the CTA's 128 x 64 tile (16
// accumulator tiles of 8 x 8 = 64 registers); 2 CTAs per SM = 4 warps per SM sub-partition.  A single warp cannot
// issue DMMA.8x8x4 back to back (ptxas pads them: ~26 cycles between two of the same warp against 16 cycles of pipe
// time each), so the FP64 pipe needs several resident warps per sub-partition to stay full -- with 2 (the first
// version: 4 warps x 64 x 32) the kernel reached 58 % of the pipe's peak.
__global__ void __launch_bounds__(256, 2) chol_update_kernel(const UpdateArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wm = warp & 3, wn = warp >> 2;
  int j0, i0;
  update_tile_of(a, blockIdx.x, j0, i0);
  if (j0 >= a.n_cols || i0 >= a.n_rows) return;   // the second half of a last, narrow block column
  const int nk = (a.kb + UP_BK - 1) / UP_BK;
  // the C tile is read-modify-written at the very end: start it on its way from HBM to L2 now (2 lines per thread)
  {
    const int col = j0 + (tid >> 2), r = i0 + (tid & 3) * 32;
    if (col < a.n_cols && r < a.n_rows) {
      const double* pc = a.A + (size_t)col * a.ld + r;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pc));
      if (r + 16 < a.n_rows) asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + 16));
    }
  }
  auto load = [&](int kc, int st) {
    double* as = sm + st * UP_STAGE_DOUBLES;
    double* bs = as + UP_BK * UP_AST;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + 256 * q;
      const int kk = idx >> 6, mm = (idx & 63) * 2;
      const int kg = kc * UP_BK + kk;
      const bool ok = kg < a.kb && i0 + mm < a.n_rows;
      cp_async16(as + kk * UP_AST + mm, ok ? a.A + (size_t)(a.k0 + kg) * a.ld + i0 + mm : a.A, ok ? 16 : 0);
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int idx = tid + 256 * q;
      const int kk = idx >> 5, nn = (idx & 31) * 2;
      const int kg = kc * UP_BK + kk;
      const bool ok = kg < a.kb && j0 + nn < a.n_rows;
      cp_async16(bs + kk * UP_BST + nn, ok ? a.A + (size_t)(a.k0 + kg) * a.ld + j0 + nn : a.A, ok ? 16 : 0);
    }
  };
#pragma unroll
  for (int s = 0; s < UP_STAGES - 1; ++s) {
    if (s < nk) load(s, s);
    cp_async_commit();
  }
  double acc[4][4][2] = {};
  const int lk = lane & 3, lr = lane >> 2;
  for (int kc = 0; kc < nk; ++kc) {
    cp_async_wait<UP_STAGES - 2>();
    __syncthreads();
    if (kc + UP_STAGES - 1 < nk) load(kc + UP_STAGES - 1, (kc + UP_STAGES - 1) % UP_STAGES);
    cp_async_commit();
    const double* as = sm + (kc % UP_STAGES) * UP_STAGE_DOUBLES + wm * 32 + lr;
    const double* bs = sm + (kc % UP_STAGES) * UP_STAGE_DOUBLES + UP_BK * UP_AST + wn * 32 + lr;
#pragma unroll
    for (int k = 0; k < UP_BK; k += 4) {
      double af[4], bf[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) af[i] = as[(k + lk) * UP_AST + 8 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = bs[(k + lk) * UP_BST + 8 * j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma(acc[i][j], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
  // C(i, j) -= acc, lower triangle only.  The tile comes from HBM (~0.8 us): 16 loads per thread are in flight
  // before the first store, so the read-modify-write latency is paid twice per tile, not once per element.
  const int ib = i0 + wm * 32 + lr, jb = j0 + wn * 32 + 2 * lk;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    double cv[4][2][2];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int col = jb + 8 * (2 * h + jj) + q;
        const double* cp = a.A + (size_t)col * a.ld;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = ib + 8 * i;
          cv[i][jj][q] = (col < a.n_cols && row >= col && row < a.n_rows) ? cp[row] : 0.0;
        }
      }
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int col = jb + 8 * (2 * h + jj) + q;
        double* cp = a.A + (size_t)col * a.ld;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = ib + 8 * i;
          if (col < a.n_cols && row >= col && row < a.n_rows) cp[row] = cv[i][jj][q] - acc[i][2 * h + jj][q];
        }
      }
  }
}

SmemOptIn g_diag_optin, g_panel_optin, g_update_optin;

// ---------------------------------------------------------------------------------------------------------
// K5d: back-substitution  L^T x = r  (r in x on entry), the last step of the reduced solve.
//
// Row i of the row-major buffer is row i of U = L^T, contiguous in j: x_i = (r_i - sum_{j>i} U(i,j) x_j) / U(i,i).
// One CTA per 128-row block, bottom block first (blockIdx 0): it streams its rows against the 128-wide chunks of x
// that the CTAs below it publish (flag per block, release / acquire at gpu scope; a CTA only ever waits for CTAs
// with a smaller blockIdx, which were dispatched before it, so the wait cannot deadlock), starting the loads of a
// chunk's tile BEFORE it waits for the chunk, then solves its own 128 x 128 triangle in four 32-row sub-steps with
// the pre-inverted 32 x 32 diagonal blocks (trsv_inv32_kernel) and publishes its part of x.
// ---------------------------------------------------------------------------------------------------------
constexpr int TS_B = 128;        // rows per CTA = columns per chunk of x
constexpr int TS_WARPS = 16;     // 8 rows per warp

// inverse of every 32 x 32 diagonal block of L: one warp per block; the block is staged in shared memory (32 coalesced
// column loads in flight at once), then lane = column runs the same recurrence as warp_potrf32
__global__ void __launch_bounds__(128) trsv_inv32_kernel(const double* __restrict__ A, int ld, int n,
                                                        double* __restrict__ invd) {
  __shared__ double blk_all[4][SB * IST];
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int b0 = b * SB;
  if (b0 >= n) return;                                 // whole warps leave together
  double* blk = blk_all[threadIdx.x >> 5];             // L(b0 + r, b0 + c) at blk[c * IST + r]
  const int w = min(SB, n - b0);                       // the last block may be narrower: padded with the identity
#pragma unroll
  for (int c = 0; c < SB; ++c)
    blk[c * IST + lane] = (lane < w && c < w) ? A[(size_t)(b0 + c) * ld + b0 + lane] : (lane == c ? 1.0 : 0.0);
  __syncwarp();
  const double rdiag = 1.0 / blk[lane * IST + lane];
  double m[SB];
#pragma unroll
  for (int i = 0; i < SB; ++i) {
    double s0 = (i == lane) ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int p = 0; p < i; ++p) {
      const double l = blk[p * IST + i];               // uniform address across the warp: broadcast
      if ((p & 3) == 0) s0 = fma(-l, m[p], s0);
      else if ((p & 3) == 1) s1 = fma(-l, m[p], s1);
      else if ((p & 3) == 2) s2 = fma(-l, m[p], s2);
      else s3 = fma(-l, m[p], s3);
    }
    m[i] = ((s0 + s1) + (s2 + s3)) * __shfl_sync(0xffffffffu, rdiag, i);
  }
#pragma unroll
  for (int i = 0; i < SB; ++i) invd[(size_t)b * SB * SB + lane * SB + i] = m[i];   // inv(n = i, k = lane) at [k * 32 + n]
}

constexpr int TS_UST = TS_B + 2;     // shared-memory row stride of the diagonal tile
constexpr int TS_IST = SB + 1;       // and of the transposed inverses (lane = column reads stay conflict-free)

__global__ void __launch_bounds__(TS_WARPS * 32, 1) chol_trsv_kernel(const double* __restrict__ A, int ld, int n,
                                                                     const double* __restrict__ invd, double* x,
                                                                     int* flags, int* err) {
  extern __shared__ __align__(16) double sm[];
  double* Us = sm;                              // [128][TS_UST]  U(r0 + i, r0 + j): the block's own triangle
  double* Is = Us + TS_B * TS_UST;              // [4][32][TS_IST] inv(L_ss)(k, i) at [s][k * TS_IST + i]
  double* rs = Is + 4 * SB * TS_IST;            // [128]
  double* xs = rs + TS_B;                       // [128]
  double* ts = xs + TS_B;                       // [32]
  const int nblk = (n + TS_B - 1) / TS_B;
  const int blk = nblk - 1 - (int)blockIdx.x;         // bottom block first
  const int r0 = blk * TS_B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // the block's own triangle and its four inverted diagonal blocks go to shared memory now (asynchronously): the
  // final sub-steps are on the critical path of every block above this one and must not wait for global memory
  for (int idx = tid; idx < TS_B * (TS_B / 2); idx += TS_WARPS * 32) {
    const int i = idx >> 6, j2 = (idx & 63) * 2;
    const bool ok = r0 + i < n && r0 + j2 < n;
    cp_async16(Us + i * TS_UST + j2, ok ? A + (size_t)(r0 + i) * ld + r0 + j2 : A, ok ? 16 : 0);
  }
  cp_async_commit();
  for (int idx = tid; idx < 4 * SB * SB; idx += TS_WARPS * 32) {
    const int s = idx >> 10, i = (idx >> 5) & 31, k = idx & 31;        // invd: inv(k, i) at [i * 32 + k]
    const int b32 = r0 / SB + s;
    Is[s * SB * TS_IST + k * TS_IST + i] = (b32 * SB < n) ? invd[(size_t)b32 * SB * SB + i * SB + k] : 0.0;
  }
  if (tid < TS_B) xs[tid] = 0.0;                      // entries past the end of the matrix stay zero
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int c = nblk - 1; c > blk; --c) {
    const int c0 = c * TS_B + 4 * lane;
    // tile rows of this warp against chunk c: issued before the chunk is waited for
    double2 t0[8], t1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = r0 + 8 * warp + i;              // row < r0 + 128 <= c * 128 <= n - 1: always a real row
      const double* p = A + (size_t)row * ld + c0;
      t0[i] = (c0 < n) ? *reinterpret_cast<const double2*>(p) : make_double2(0.0, 0.0);
      t1[i] = (c0 + 2 < n) ? *reinterpret_cast<const double2*>(p + 2) : make_double2(0.0, 0.0);
      if (c0 + 1 >= n) t0[i].y = 0.0;                 // columns >= n of the buffer are not part of U
      if (c0 + 3 >= n) t1[i].y = 0.0;
    }
    if (lane == 0) {
      int spins = 0;
      int f;
      do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(flags + c) : "memory");
        if (!f && ++spins > (1 << 26)) { atomicExch(err, 1); break; }    // never hang the device on a lost flag
      } while (!f);
    }
    __syncwarp();
    double xv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) xv[q] = (c0 + q < n) ? __ldcg(x + c0 + q) : 0.0;     // written by another SM: L2 only
#pragma unroll
    for (int i = 0; i < 8; ++i)
      acc[i] = fma(t0[i].x, xv[0], fma(t0[i].y, xv[1], fma(t1[i].x, xv[2], fma(t1[i].y, xv[3], acc[i]))));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    const int row = r0 + 8 * warp + i;
    if (lane == 0) rs[8 * warp + i] = (row < n) ? x[row] - acc[i] : 0.0;
  }
  cp_async_wait<0>();
  __syncthreads();
  // own triangle, bottom sub-block first, everything from shared memory
  for (int s = TS_B / SB - 1; s >= 0; --s) {
    if (r0 + s * SB >= n) continue;                    // block-uniform
    // t_i = rs_i - sum over the later sub-blocks of this block: warps 0..3 take 8 rows each, lanes the columns
    if (warp < 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int li = s * SB + 8 * warp + i;
        double v = 0.0;
#pragma unroll
        for (int jj = 0; jj < 3; ++jj) {
          const int j = (s + 1) * SB + lane + 32 * jj;
          if (j < TS_B) v = fma(Us[li * TS_UST + j], xs[j], v);      // xs of sub-blocks not yet solved is never read:
        }                                                           // j >= (s+1)*32 were solved in earlier sub-steps
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) ts[8 * warp + i] = rs[li] - v;
      }
    }
    __syncthreads();
    // x_i = sum_{k >= i} inv(L_ss)(k, i) t_k
    if (warp == 0) {
      const double* inv = Is + s * SB * TS_IST + lane;
      double v0 = 0.0, v1 = 0.0;
#pragma unroll
      for (int k = 0; k < SB; k += 2) {               // inv(k, i) = 0 for k < i: no branch needed
        v0 = fma(inv[k * TS_IST], ts[k], v0);
        v1 = fma(inv[(k + 1) * TS_IST], ts[k + 1], v1);
      }
      xs[s * SB + lane] = v0 + v1;
    }
    __syncthreads();
  }
  if (tid < TS_B && r0 + tid < n) x[r0 + tid] = xs[tid];
  __threadfence();
  __syncthreads();
  if (tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + blk), "r"(1) : "memory");
}

SmemOptIn g_trsv_optin;

}  // namespace

size_t chol_workspace_doubles() { return (size_t)NB * NB + 4 * SB * SB; }

void chol_diag(double* A, int ld, int k0, int kb, double* ws, int* info, cudaStream_t s) {
  const size_t smem = (size_t)(NB * LST + SB * IST) * sizeof(double);
  g_diag_optin.ensure(chol_diag_kernel, smem);
  chol_diag_kernel<<<1, 128, smem, s>>>(A, ld, k0, kb, ws, ws + NB * NB, info);
  RCC_CUDA(cudaGetLastError());
}

void chol_panel(double* A, int ld, int n_rows, int k0, int kb, const double* ws, cudaStream_t s) {
  const int row0 = k0 + kb;
  if (row0 >= n_rows) return;
  const size_t smem = (size_t)(NB * XST + 10 * SB * IST) * sizeof(double);
  g_panel_optin.ensure(chol_panel_kernel, smem);
  chol_panel_kernel<<<ceil_div(n_rows - (row0 & ~1), 64), 128, smem, s>>>(A, ld, n_rows, k0, kb, row0, ws, ws + NB * NB);
  RCC_CUDA(cudaGetLastError());
}

size_t chol_trsv_workspace_doubles(int n) { return (size_t)((n + SB - 1) / SB) * SB * SB; }
int chol_trsv_flags(int n) { return (n + TS_B - 1) / TS_B + 1; }   // + 1: the error flag

void chol_trsv(const double* A, int ld, int n, double* x, double* invd, int* flags, cudaStream_t s) {
  const int nblk = (n + TS_B - 1) / TS_B;
  RCC_CUDA(cudaMemsetAsync(flags, 0, (size_t)(nblk + 1) * sizeof(int), s));
  trsv_inv32_kernel<<<ceil_div((int64_t)((n + SB - 1) / SB) * 32, 128), 128, 0, s>>>(A, ld, n, invd);
  RCC_CUDA(cudaGetLastError());
  const size_t smem = (size_t)(TS_B * TS_UST + 4 * SB * TS_IST + 2 * TS_B + SB) * sizeof(double);
  g_trsv_optin.ensure(chol_trsv_kernel, smem);
  chol_trsv_kernel<<<nblk, TS_WARPS * 32, smem, s>>>(A, ld, n, invd, x, flags, flags + nblk);
  RCC_CUDA(cudaGetLastError());
}

void chol_update(double* A, int ld, int n_rows, int n_cols, int k0, int kb, int first_blk, int blk_stride, int n_blks,
                 cudaStream_t s) {
  if (n_blks <= 0 || first_blk * NB >= n_cols) return;
  static_assert(NB == UP_BM, "row tiles and block columns share one index");
  const int tm = ceil_div(n_rows, UP_BM);
  UpdateArgs a{A, ld, n_rows, n_cols, k0, kb, first_blk, blk_stride, tm};
  const size_t smem = (size_t)UP_STAGES * UP_STAGE_DOUBLES * sizeof(double);
  g_update_optin.ensure(chol_update_kernel, smem);
  // lower-triangular tiles only: block column b has tm - (first_blk + b * stride) row tiles, two halves each
  const int64_t n_ctas = 2 * ((int64_t)n_blks * (tm - first_blk) - (int64_t)blk_stride * n_blks * (n_blks - 1) / 2);
  if (n_ctas <= 0) return;
  chol_update_kernel<<<(unsigned)n_ctas, 256, smem, s>>>(a);
  RCC_CUDA(cudaGetLastError());
}

}  // namespace rcc

// ---------------------------------------------------------------------------------------------------------
// panel driver
// ---------------------------------------------------------------------------------------------------------
namespace rcc {

void CholDriver::init() {
  if (panel_stream) return;
  int lo = 0, hi = 0;
  RCC_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // hi = numerically lowest = highest priority
  RCC_CUDA(cudaStreamCreateWithPriority(&panel_stream, cudaStreamNonBlocking, hi));
  for (auto& e : ev_panel) RCC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : ev_col) RCC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  RCC_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
  RCC_CUDA(cudaMalloc(&ws, chol_workspace_doubles() * sizeof(double)));
  RCC_CUDA(cudaMalloc(&info, sizeof(int)));
}

void CholDriver::destroy() {
  if (panel_stream) {
    cudaStreamSynchronize(panel_stream);
    cudaStreamDestroy(panel_stream);
  }
  for (auto e : ev_panel)
    if (e) cudaEventDestroy(e);
  for (auto e : ev_col)
    if (e) cudaEventDestroy(e);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ws) cudaFree(ws);
  if (stage) cudaFree(stage);
  if (info) cudaFree(info);
  *this = CholDriver();
}

void chol_factor(double* A, int n, int ld, int n_rows, int rank, int n_ranks, ncclComm_t comm, cudaStream_t main,
                 CholDriver& d) {
  d.init();
  RCC_REQUIRE((ld & 1) == 0 && ld >= n_rows + (n_rows & 1), RCC_BAD_ARG,
              "dense Cholesky: the leading dimension must be even and cover the rows in pairs");
  if (n_ranks <= 1 || comm == nullptr) {
    n_ranks = 1;
    rank = 0;
  }
  const int nblk = (n + CHOL_NB - 1) / CHOL_NB;
  cudaStream_t ps = d.panel_stream;
  if (n_ranks > 1 && d.stage_cap < (size_t)n_rows * CHOL_NB) {
    if (d.stage) RCC_CUDA(cudaFree(d.stage));
    d.stage = nullptr;
    d.stage_cap = (size_t)n_rows * CHOL_NB;
    RCC_CUDA(cudaMalloc(&d.stage, d.stage_cap * sizeof(double)));
  }
  RCC_CUDA(cudaEventRecord(d.ev_fork, main));
  RCC_CUDA(cudaStreamWaitEvent(ps, d.ev_fork, 0));
  RCC_CUDA(cudaMemsetAsync(d.info, 0, sizeof(int), ps));
  for (int K = 0; K < nblk; ++K) {
    const int k0 = K * CHOL_NB, kb = std::min(CHOL_NB, n - k0);
    const int owner = K % n_ranks;
    if (owner == rank) {
      // block column K is complete once panel K-1 has been applied to it (look-ahead update on `main`)
      if (K > 0) RCC_CUDA(cudaStreamWaitEvent(ps, d.ev_col[K & 3], 0));
      chol_diag(A, ld, k0, kb, d.ws, d.info, ps);
      chol_panel(A, ld, n_rows, k0, kb, d.ws, ps);
      d.launches += 2;
    }
    if (n_ranks > 1) {
      // The panel = rows k0 .. k0+kb-1 of the row-major buffer from column k0 on.  In place that is one contiguous
      // range only together with the unused strict upper part of the rows in between (30 MB per panel whatever
      // k0 is); packed through a staging buffer it is kb x (n_rows - k0): half the bytes on average, and the late
      // panels -- the ones on the critical path of the ranks -- become latency-sized.
      const size_t width = (size_t)(n_rows - k0);
      double* p = A + (size_t)k0 * ld + k0;
      if (owner == rank)
        RCC_CUDA(cudaMemcpy2DAsync(d.stage, width * sizeof(double), p, (size_t)ld * sizeof(double),
                                   width * sizeof(double), kb, cudaMemcpyDeviceToDevice, ps));
      ncclResult_t r = ncclBroadcast(d.stage, d.stage, width * kb, ncclDouble, owner, comm, ps);
      if (r != ncclSuccess) throw Error(RCC_NCCL_ERROR, std::string("ncclBroadcast(panel): ") + ncclGetErrorString(r));
      if (owner != rank)
        RCC_CUDA(cudaMemcpy2DAsync(p, (size_t)ld * sizeof(double), d.stage, width * sizeof(double),
                                   width * sizeof(double), kb, cudaMemcpyDeviceToDevice, ps));
    }
    RCC_CUDA(cudaEventRecord(d.ev_panel[K & 3], ps));
    RCC_CUDA(cudaStreamWaitEvent(main, d.ev_panel[K & 3], 0));
    if (K + 1 < nblk) {
      if ((K + 1) % n_ranks == rank) {
        chol_update(A, ld, n_rows, n, k0, kb, K + 1, 1, 1, main);     // look-ahead: the next panel's block column first
        RCC_CUDA(cudaEventRecord(d.ev_col[(K + 1) & 3], main));
        d.launches += 1;
      }
      int first = K + 2;
      first += ((rank - first) % n_ranks + n_ranks) % n_ranks;          // smallest owned block column >= K + 2
      if (first < nblk) {
        chol_update(A, ld, n_rows, n, k0, kb, first, n_ranks, (nblk - 1 - first) / n_ranks + 1, main);
        d.launches += 1;
      }
    }
  }
}

}  // namespace rcc
