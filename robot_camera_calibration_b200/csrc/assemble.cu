// K2 -- fused reprojection residual + analytic Jacobian + Gauss-Newton normal
// equation assembly.  The Jacobian never reaches HBM.
//
// Replaces: the Ceres cost-functor evaluation and the
// linearisation / J^T J builder that the reference's missing optimiser stage
// would run between camera_pose.cpp (initial guesses) and opt_visualization.cpp.
//
// Work decomposition.  Observation blocks are sorted by the 6-dof block that
// "owns" the pass (the eliminated set in the E pass, the kept set in the F
// pass) and cut into chunks that share (own block, camera).  One warp walks one
// chunk.  Inside the warp a thread-group of TPB lanes serves one observation
// block:
//   phase 1  lanes 0..3 of the group evaluate one tag corner each (2 residual
//            rows, all Jacobian column groups) and stage the rows in shared
//            memory;
//   phase 2  every lane of the group owns one 6x6 tile I^T J of the block's
//            8 x NCOL row matrix and accumulates it in 36 FP64 registers across
//            all blocks of the chunk.
// The per-block cross tile J_own^T J_other (the Schur off-diagonal block W) is
// written straight to HBM; everything else leaves the warp once per chunk as a
// TPB x 36 partial that finalize_* reduce in a fixed order (deterministic, no
// FP64 atomics).
#include "common.cuh"
#include "kernels.h"
#include "model.cuh"

namespace rcc {

// stores the column groups of one corner's two residual rows into the staged row layout
// [O(6) | T(6) | S1(6) | S2: p1 p2 k3 r 0 0 | X(6)]
template <bool RIG, bool OWN_IS_VIEW>
struct RowSink {
  double* row[2];
  double w;     // row scale sqrt(rho'(s)) of the block (1 for the trivial loss)
  double c22;   // sqrt(rho(s)) on the block's first row, 0 elsewhere: S2S2[4][4] sums the robust cost
  __device__ __forceinline__ void put6(int i, int col, const double* v) {
    double2* d = reinterpret_cast<double2*>(row[i] + col);
    d[0] = make_double2(w * v[0], w * v[1]);
    d[1] = make_double2(w * v[2], w * v[3]);
    d[2] = make_double2(w * v[4], w * v[5]);
  }
  __device__ __forceinline__ void shared(int i, const double* js, double r) {
    put6(i, 12, js);
    double2* d = reinterpret_cast<double2*>(row[i] + 18);
    d[0] = make_double2(w * js[6], w * js[7]);
    d[1] = make_double2(w * js[8], w * r);
    d[2] = make_double2(i == 0 ? c22 : 0.0, 0.0);
  }
  __device__ __forceinline__ void marker(int i, const double* jm) { put6(i, OWN_IS_VIEW ? 6 : 0, jm); }
  __device__ __forceinline__ void view(int i, const double* jv) { put6(i, OWN_IS_VIEW ? 0 : 6, jv); }
  __device__ __forceinline__ void ext(int i, const double* jx) { put6(i, 24, jx); }
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

#ifndef RCC_K2_MIN_CTAS
#define RCC_K2_MIN_CTAS 2
#endif

// per-warp shared memory (doubles): staged rows | 2 x BPW other-pose records | own pose, ext pose, shared params
template <bool RIG>
struct WarpSmem {
  using PG = PassGeom<RIG>;
  static constexpr int ROWS = 0;
  static constexpr int STAGE = PG::WARP_SMEM;
  static constexpr int CONSTS = STAGE + 2 * PG::BPW * POSEX;
  static constexpr int TOTAL = CONSTS + 2 * POSEX + 16;
};

template <bool RIG, bool EPASS, bool OWN_IS_VIEW, bool LOSS>
__global__ void __launch_bounds__(PassGeom<RIG>::WARPS * 32, RCC_K2_MIN_CTAS)
assemble_kernel(const AssembleArgs a) {
  using PG = PassGeom<RIG>;
  using WS = WarpSmem<RIG>;
  constexpr int TPB = PG::TPB, BPW = PG::BPW, NCOL = PG::RS, BS = PG::BLK_STRIDE, RSTR = PG::RED_STRIDE;
  extern __shared__ __align__(16) double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk_id = blockIdx.x * PG::WARPS + warp;
  if (chunk_id >= a.n_chunks) return;  // whole warp leaves together
  double* wsm = smem + warp * WS::TOTAL;
  double* rows = wsm + WS::ROWS;
  double* stage = wsm + WS::STAGE;    // [2][BPW][POSEX] expanded pose of the other block, double buffered
  double* own_x = wsm + WS::CONSTS;   // [POSEX] expanded pose of the own block
  double* ext_x = own_x + POSEX;      // [POSEX] body_T_cam (rig)
  double* sh = ext_x + POSEX;         // [SP] shared parameters of the chunk's camera

  const Chunk ch = a.chunks[chunk_id];
  const int b = lane / TPB;
  const int t = lane - b * TPB;
  const int gi_id = tile_I(RIG, EPASS, t), gj_id = tile_J(RIG, EPASS, t);
  const bool lane_on = (lane < BPW * TPB) && (gi_id != G_NONE);
  const int gi = gi_id * 6, gj = gj_id * 6;
  const bool eval_lane = (lane < BPW * TPB) && (t < 4);
  const double* oth_table = OWN_IS_VIEW ? a.marker_x : a.view_x;

  // chunk constants -> shared memory
  {
    const double* src = (OWN_IS_VIEW ? a.view_x : a.marker_x) + (size_t)ch.own * POSEX;
    if (lane < POSEX) own_x[lane] = src[lane];
    if (RIG && lane < POSEX) ext_x[lane] = a.ext_x[(size_t)ch.cam * POSEX + lane];
    if (lane < PG::SP) sh[lane] = a.shared[ch.cam * PG::SP + lane];
  }
  // software pipeline: other-block indices two iterations ahead (lane q holds block q),
  // other-block pose records one iteration ahead (cp.async into stage[buf])
  auto load_oth = [&](int it) -> int {
    return (lane < BPW && it + lane < ch.count) ? a.oth[(int64_t)ch.start + it + lane] : 0;
  };
  auto issue_stage = [&](int it, int buf, int oth_reg) {
    // BPW records x 12 16-byte pieces
    constexpr int PIECES = BPW * (POSEX / 2);
#pragma unroll
    for (int r = 0; r < (PIECES + 31) / 32; ++r) {   // uniform trip count: the shuffle needs every lane
      const int idx = lane + 32 * r;
      const int q = min(idx / (POSEX / 2), BPW - 1), c = idx - q * (POSEX / 2);
      const int o = __shfl_sync(0xffffffffu, oth_reg, q);
      if (idx < PIECES && it + q < ch.count)
        cp_async16(stage + (buf * BPW + q) * POSEX + 2 * c, oth_table + (size_t)o * POSEX + 2 * c);
    }
    cp_async_commit();
  };
  int oth_cur = load_oth(0);
  int oth_nxt = load_oth(BPW);
  issue_stage(0, 0, oth_cur);
  double2 px_nxt = make_double2(0.0, 0.0);
  if (eval_lane && b < ch.count) px_nxt = *reinterpret_cast<const double2*>(a.pix + ((int64_t)ch.start + b) * 8 + 2 * t);

  double acc[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0.0;
  double* blk = rows + b * BS;
  int buf = 0;

  for (int it = 0; it < ch.count; it += BPW, buf ^= 1) {
    const bool valid = (it + b) < ch.count;
    const int64_t g = (int64_t)ch.start + it + b;
    cp_async_wait_all();
    __syncwarp();
    // prefetch for the next iterations (overlaps with this iteration's arithmetic)
    const double2 px = px_nxt;
    if (it + BPW < ch.count) {
      issue_stage(it + BPW, buf ^ 1, oth_nxt);
      oth_nxt = load_oth(it + 2 * BPW);
      if (eval_lane && it + BPW + b < ch.count)
        px_nxt = *reinterpret_cast<const double2*>(a.pix + (g + BPW) * 8 + 2 * t);
    }
    // ---- phase 1: corner evaluation ---------------------------------------
    const double* ox_rec = stage + (buf * BPW + min(b, BPW - 1)) * POSEX;
    const double* vx = OWN_IS_VIEW ? own_x : ox_rec;
    const double* mx = OWN_IS_VIEW ? ox_rec : own_x;
    BlockGeom<RIG> geo;
    double ox = 0.0, oy = 0.0;
    if (valid && eval_lane) {
      block_geometry<RIG>(vx, mx, RIG ? ext_x : nullptr, geo);
      corner_xy(t, mx[PX_HS], ox, oy);
    }
    double w = 1.0, c22 = 0.0;
    if (LOSS) {
      // s = sum of the 8 squared residuals of the block: residual-only evaluation, then a
      // 4-lane gather over the block's corner lanes (every lane of the warp takes part)
      double s_own = 0.0;
      if (valid && eval_lane) {
        CornerRows<RIG> c;
        eval_corner<RIG, false>(geo, sh, ox, oy, px.x, px.y, c);
        s_own = c.r[0] * c.r[0] + c.r[1] * c.r[1];
      }
      const int l0 = min(b, BPW - 1) * TPB;
      double s_blk = __shfl_sync(0xffffffffu, s_own, l0);
      s_blk += __shfl_sync(0xffffffffu, s_own, l0 + 1);
      s_blk += __shfl_sync(0xffffffffu, s_own, l0 + 2);
      s_blk += __shfl_sync(0xffffffffu, s_own, l0 + 3);
      double rho1;
      const double rho = robust_rho(a.loss, a.loss_a2, s_blk, rho1);
      w = sqrt(rho1);
      c22 = (t == 0) ? sqrt(fmax(rho, 0.0)) : 0.0;
    }
    if (valid && eval_lane) {
      RowSink<RIG, OWN_IS_VIEW> sink{{blk + (2 * t) * NCOL, blk + (2 * t + 1) * NCOL}, w, c22};
      double r0, r1;
      const double depth = eval_corner_emit<RIG>(geo, sh, ox, oy, px.x, px.y, sink, r0, r1);
      if (!(depth > 0.0) || !isfinite(r0) || !isfinite(r1)) *a.fail_flag = 1;
    }
    __syncwarp();
    // ---- phase 2: 6x6 tile  I^T J  over the block's 8 rows ----------------
    if (valid && lane_on) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double2* ri = reinterpret_cast<const double2*>(blk + k * NCOL + gi);
        const double2* rj = reinterpret_cast<const double2*>(blk + k * NCOL + gj);
        const double2 i0 = ri[0], i1 = ri[1], i2 = ri[2];
        const double2 j0 = rj[0], j1 = rj[1], j2 = rj[2];
        const double I[6] = {i0.x, i0.y, i1.x, i1.y, i2.x, i2.y};
        const double J[6] = {j0.x, j0.y, j1.x, j1.y, j2.x, j2.y};
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 6; ++j) acc[i * 6 + j] = fma(I[i], J[j], acc[i * 6 + j]);
      }
      if (EPASS && t == 1) {
        // cross block W = J_own^T J_other of this observation block
        double2* w = reinterpret_cast<double2*>(a.W + g * 36);
#pragma unroll
        for (int i = 0; i < 18; ++i) {
          w[i] = make_double2(acc[2 * i], acc[2 * i + 1]);
          acc[2 * i] = 0.0;
          acc[2 * i + 1] = 0.0;
        }
      }
    }
    __syncwarp();
  }

  // ---- chunk epilogue: sum the BPW thread-groups, one partial per tile ------
#pragma unroll
  for (int k = 0; k < 36; ++k) rows[lane * RSTR + k] = acc[k];
  __syncwarp();
  double* out = a.partials + (size_t)chunk_id * (TPB * 36);
  for (int o = lane; o < TPB * 36; o += 32) {
    const int tt = o / 36, k = o - tt * 36;
    double s = 0.0;
#pragma unroll
    for (int bb = 0; bb < BPW; ++bb) s += rows[(bb * TPB + tt) * RSTR + k];
    out[o] = s;
  }
}

template <bool RIG, bool EPASS, bool OWN_IS_VIEW, bool LOSS>
static void launch_assemble_t(const AssembleArgs& a, cudaStream_t s) {
  using PG = PassGeom<RIG>;
  if (a.n_chunks == 0) return;
  const size_t smem = PG::WARPS * WarpSmem<RIG>::TOTAL * sizeof(double);
  auto k = assemble_kernel<RIG, EPASS, OWN_IS_VIEW, LOSS>;
  static bool attr_set = false;
  if (!attr_set) {
    RCC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int grid = ceil_div(a.n_chunks, PG::WARPS);
  k<<<grid, PG::WARPS * 32, smem, s>>>(a);
  RCC_CUDA(cudaGetLastError());
}

template <bool LOSS>
static void launch_assemble_l(bool rig, bool epass, bool own_is_view, const AssembleArgs& a, cudaStream_t s) {
  const int sel = (rig ? 4 : 0) | (epass ? 2 : 0) | (own_is_view ? 1 : 0);
  switch (sel) {
    case 0: launch_assemble_t<false, false, false, LOSS>(a, s); break;
    case 1: launch_assemble_t<false, false, true, LOSS>(a, s); break;
    case 2: launch_assemble_t<false, true, false, LOSS>(a, s); break;
    case 3: launch_assemble_t<false, true, true, LOSS>(a, s); break;
    case 4: launch_assemble_t<true, false, false, LOSS>(a, s); break;
    case 5: launch_assemble_t<true, false, true, LOSS>(a, s); break;
    case 6: launch_assemble_t<true, true, false, LOSS>(a, s); break;
    default: launch_assemble_t<true, true, true, LOSS>(a, s); break;
  }
}

void launch_assemble(bool rig, bool epass, bool own_is_view, const AssembleArgs& a, cudaStream_t s) {
  if (a.loss != 0) launch_assemble_l<true>(rig, epass, own_is_view, a, s);
  else launch_assemble_l<false>(rig, epass, own_is_view, a, s);
}

// ---------------------------------------------------------------------------
// finalize: one warp per own block sums its chunk partials in chunk order.
// ---------------------------------------------------------------------------
template <bool RIG, bool EPASS>
__global__ void __launch_bounds__(160) finalize_side_kernel(const FinalizeSideArgs a) {
  // one CTA per own block; thread (q, k) sums entry k of tile q over the block's chunks.
  // q: 0 = O x O, 1 = O x S1, 2 = O x S2 (col 3 = gradient), 3 = O x X (rig)
  using PG = PassGeom<RIG>;
  constexpr int TPB = PG::TPB, SP = PG::SP, NQ = RIG ? 4 : 3;
  constexpr int T_S1 = EPASS ? 2 : 1;
  const int i = blockIdx.x;
  const int tid = threadIdx.x;
  double* hos = a.Hos + (size_t)i * 6 * a.n_shared;
  for (int k = tid; k < 6 * a.n_shared; k += blockDim.x) hos[k] = 0.0;
  __syncthreads();
  const int q = tid / 36, k = tid - 36 * q;
  if (q >= NQ) return;
  const int tile = (q == 0) ? 0 : T_S1 + (q - 1);
  const int r = k / 6, col = k - 6 * r;
  const int c0 = a.chunk_ptr[i], c1 = a.chunk_ptr[i + 1];
  double total = 0.0, per_cam = 0.0;
  int cam = (c0 < c1) ? a.chunks[c0].cam : 0;
  auto flush = [&](int cm) {
    double* hrow = hos + r * a.n_shared + cm * SP;
    if (q == 1) hrow[col] = per_cam;
    else if (q == 2 && col < 3) hrow[6 + col] = per_cam;
    else if (q == 3) hrow[9 + col] = per_cam;
    per_cam = 0.0;
  };
  for (int c = c0; c < c1; ++c) {
    const int cm = a.chunks[c].cam;
    if (cm != cam) {
      flush(cam);
      cam = cm;
    }
    const double v = a.partials[(size_t)c * (TPB * 36) + tile * 36 + k];
    total += v;
    per_cam += v;
  }
  if (c0 < c1) flush(cam);
  if (q == 0) a.Hoo[(size_t)i * 36 + k] = total;
  if (q == 2 && col == 3) a.go[(size_t)i * 6 + r] = total;
}

void launch_finalize_side(bool rig, bool epass, const FinalizeSideArgs& a, cudaStream_t s) {
  if (a.n_own == 0) return;
  const int grid = a.n_own;
  if (rig) {
    if (epass) finalize_side_kernel<true, true><<<grid, 160, 0, s>>>(a);
    else finalize_side_kernel<true, false><<<grid, 160, 0, s>>>(a);
  } else {
    if (epass) finalize_side_kernel<false, true><<<grid, 160, 0, s>>>(a);
    else finalize_side_kernel<false, false><<<grid, 160, 0, s>>>(a);
  }
  RCC_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// finalize_shared: shared x shared tiles, gradient and cost per camera.
// Stage 1: CTA (slice, camera) sums its share of the camera's chunk list for
// every shared tile (fixed order).  Stage 2: one CTA per camera sums the
// FIN_SLICES partials in order and scatters them into H_ss / g_s / cost.
// ---------------------------------------------------------------------------
struct SharedTile { int from_f; int tile; };
template <bool RIG>
__device__ __forceinline__ SharedTile shared_tile(int q) {
  // q: 0 S1S1 (E)  1 S1S2 (F)  2 S2S2 (F)  3 S1X (E)  4 XX (E)  5 S2X (F)
  switch (q) {
    case 0: return {0, RIG ? 5 : 4};
    case 1: return {1, RIG ? 4 : 3};
    case 2: return {1, RIG ? 5 : 4};
    case 3: return {0, 6};
    case 4: return {0, 7};
    default: return {1, 6};
  }
}

template <bool RIG>
__global__ void __launch_bounds__(256) finalize_shared_partial_kernel(const FinalizeSharedArgs a) {
  constexpr int TPB = PassGeom<RIG>::TPB, NQ = RIG ? 6 : 3;
  __shared__ double red[7 * 36];
  const int slice = blockIdx.x, cam = blockIdx.y;
  const int tid = threadIdx.x;
  const int sl = tid / 36, k = tid - 36 * sl;
  for (int q = 0; q < NQ; ++q) {
    const SharedTile st = shared_tile<RIG>(q);
    const double* part = st.from_f ? a.part_f : a.part_e;
    const int32_t* list = st.from_f ? a.cam_chunks_f + a.cam_ptr_f[cam] : a.cam_chunks_e + a.cam_ptr_e[cam];
    const int n = st.from_f ? a.cam_ptr_f[cam + 1] - a.cam_ptr_f[cam] : a.cam_ptr_e[cam + 1] - a.cam_ptr_e[cam];
    const int per = (n + FIN_SLICES - 1) / FIN_SLICES;
    const int lo = slice * per, hi = min(n, lo + per);
    if (sl < 7) {
      double s = 0.0;
      for (int c = lo + sl; c < hi; c += 7) s += part[(size_t)list[c] * (TPB * 36) + st.tile * 36 + k];
      red[sl * 36 + k] = s;
    }
    __syncthreads();
    if (tid < 36) {
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < 7; ++w) s += red[w * 36 + tid];
      a.scratch[(((size_t)cam * FIN_SLICES + slice) * 6 + q) * 36 + tid] = s;
    }
    __syncthreads();
  }
}

template <bool RIG>
__global__ void __launch_bounds__(256) finalize_shared_final_kernel(const FinalizeSharedArgs a) {
  constexpr int SP = PassGeom<RIG>::SP, NQ = RIG ? 6 : 3;
  __shared__ double tile[6][36];
  const int cam = blockIdx.x;
  const int tid = threadIdx.x;
  const int ns = a.n_shared;
  double* H = a.Hss;
  const int base = cam * SP;
  for (int k = tid; k < SP * ns; k += blockDim.x) H[(size_t)base * ns + k] = 0.0;
  if (tid < NQ * 36) {
    const int q = tid / 36, k = tid - 36 * q;
    double s = 0.0;
    for (int sl = 0; sl < FIN_SLICES; ++sl) s += a.scratch[(((size_t)cam * FIN_SLICES + sl) * 6 + q) * 36 + k];
    tile[q][k] = s;
  }
  __syncthreads();
  if (tid >= 36) return;
  const int r = tid / 6, q = tid - 6 * r;
  // S1 x S1
  H[(size_t)(base + r) * ns + base + q] = tile[0][tid];
  // S1 x S2: cols 0..2 -> p1 p2 k3, col 3 -> gradient of S1
  if (q < 3) {
    H[(size_t)(base + r) * ns + base + 6 + q] = tile[1][tid];
    H[(size_t)(base + 6 + q) * ns + base + r] = tile[1][tid];
  } else if (q == 3) {
    a.gs[base + r] = tile[1][tid];
  }
  // S2 x S2
  if (r < 3 && q < 3) H[(size_t)(base + 6 + r) * ns + base + 6 + q] = tile[2][tid];
  else if (r < 3 && q == 3) a.gs[base + 6 + r] = tile[2][tid];
  else if (!a.robust && r == 3 && q == 3) a.cost2_cam[cam] = tile[2][tid];
  else if (a.robust && r == 4 && q == 4) a.cost2_cam[cam] = tile[2][tid];
  if (RIG) {
    // S1 x X
    H[(size_t)(base + r) * ns + base + 9 + q] = tile[3][tid];
    H[(size_t)(base + 9 + q) * ns + base + r] = tile[3][tid];
    // X x X
    H[(size_t)(base + 9 + r) * ns + base + 9 + q] = tile[4][tid];
    // S2 x X: rows 0..2 -> p1 p2 k3, row 3 -> gradient of X
    if (r < 3) {
      H[(size_t)(base + 6 + r) * ns + base + 9 + q] = tile[5][tid];
      H[(size_t)(base + 9 + q) * ns + base + 6 + r] = tile[5][tid];
    } else if (r == 3) {
      a.gs[base + 9 + q] = tile[5][tid];
    }
  }
}

void launch_finalize_shared(bool rig, const FinalizeSharedArgs& a, cudaStream_t s) {
  dim3 grid(FIN_SLICES, a.n_cam);
  if (rig) {
    finalize_shared_partial_kernel<true><<<grid, 256, 0, s>>>(a);
    finalize_shared_final_kernel<true><<<a.n_cam, 256, 0, s>>>(a);
  } else {
    finalize_shared_partial_kernel<false><<<grid, 256, 0, s>>>(a);
    finalize_shared_final_kernel<false><<<a.n_cam, 256, 0, s>>>(a);
  }
  RCC_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
__global__ void expand_poses_kernel(const double* __restrict__ views, int n_views, double* __restrict__ view_x,
                                    const double* __restrict__ markers, const double* __restrict__ sizes,
                                    int n_markers, double* __restrict__ marker_x,
                                    const double* __restrict__ shared, int n_cam, int sp, double* __restrict__ ext_x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_views) {
    expand_pose(views + (size_t)i * 6, view_x + (size_t)i * POSEX);
  } else if (i < n_views + n_markers) {
    const int j = i - n_views;
    expand_pose(markers + (size_t)j * 6, marker_x + (size_t)j * POSEX);
    marker_x[(size_t)j * POSEX + PX_HS] = 0.5 * sizes[j];  // half tag size rides in the record's padding
  } else if (i < n_views + n_markers + n_cam && sp == 15) {
    const int j = i - n_views - n_markers;
    expand_pose(shared + (size_t)j * sp + 9, ext_x + (size_t)j * POSEX);
  }
}

void launch_expand_poses(const double* views, int n_views, double* view_x, const double* markers,
                         const double* sizes, int n_markers, double* marker_x, const double* shared, int n_cam, int sp,
                         double* ext_x, cudaStream_t s) {
  const int n = n_views + n_markers + n_cam;
  expand_poses_kernel<<<ceil_div(n, 128), 128, 0, s>>>(views, n_views, view_x, markers, sizes, n_markers, marker_x, shared,
                                                       n_cam, sp, ext_x);
  RCC_CUDA(cudaGetLastError());
}

}  // namespace rcc
