// Dense Cholesky of the reduced system (dense.cu): kernels + the multi-stream / multi-rank panel driver.
// Internal interface, not part of the C ABI.
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

namespace rcc {

constexpr int CHOL_NB = 128;   // panel width = block-column granularity of the rank ownership

// A: lower triangle, column-major, leading dimension ld (== the row-major upper triangle of the reduced buffer).
// ws: chol_workspace_doubles() doubles (padded L_kk + the inverses of its four 32 x 32 diagonal sub-blocks).
size_t chol_workspace_doubles();
// factor the kb x kb diagonal block at k0 in place; *info (device, zero before the first panel) receives 1 + the
// first column whose pivot is not positive
void chol_diag(double* A, int ld, int k0, int kb, double* ws, int* info, cudaStream_t s);
// rows k0+kb .. n_rows-1 of the panel:  X <- X L_kk^-T
void chol_panel(double* A, int ld, int n_rows, int k0, int kb, const double* ws, cudaStream_t s);
// trailing update with panel [k0, k0+kb) of the block columns first_blk, first_blk + blk_stride, ... (n_blks of
// them, 128 columns each, columns >= n_cols excluded), rows from the diagonal down to n_rows-1
void chol_update(double* A, int ld, int n_rows, int n_cols, int k0, int kb, int first_blk, int blk_stride, int n_blks,
                 cudaStream_t s);

// back-substitution L^T x = r in place (x holds r on entry).  invd: chol_trsv_workspace_doubles(n) doubles (the
// inverses of the 32 x 32 diagonal blocks, recomputed here from whatever factor is in A); flags: chol_trsv_flags(n)
// ints, the last one is set if a wait for another CTA's part of x gave up (never expected).
size_t chol_trsv_workspace_doubles(int n);
int chol_trsv_flags(int n);
void chol_trsv(const double* A, int ld, int n, double* x, double* invd, int* flags, cudaStream_t s);

// per-handle resources of the panel pipeline
struct CholDriver {
  cudaStream_t panel_stream = nullptr;   // high priority: diagonal block, panel rows, panel broadcast
  cudaEvent_t ev_panel[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_col[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr;
  double* ws = nullptr;
  double* stage = nullptr;               // packed panel of the broadcast (kb x rows-below-and-including-the-diagonal)
  size_t stage_cap = 0;                  // doubles
  int* info = nullptr;                   // device
  int64_t launches = 0;                  // kernels launched so far (for rcc_ba_launch_count)
  void init();
  void destroy();
};

// Right-looking blocked Cholesky of the n x n matrix A (+ rows n .. n_rows-1 riding along), in place, with
// look-ahead: the panel stream factors panel K+1 while `main` applies panel K to the rest of the trailing matrix.
// n_ranks > 1: block column J belongs to rank J % n_ranks; only the owner factors and updates it, every finished
// panel is broadcast in place (ncclBroadcast on the panel stream), so that at the end every rank holds the whole
// factor.  Every rank must hold the same A on entry (the all-reduced system).  All work is ordered after what is
// already queued on `main`, and `main` has joined the panel stream on return.
void chol_factor(double* A, int n, int ld, int n_rows, int rank, int n_ranks, ncclComm_t comm, cudaStream_t main,
                 CholDriver& d);

}  // namespace rcc
