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

template <bool RIG, bool EPASS, bool OWN_IS_VIEW>
__global__ void __launch_bounds__(PassGeom<RIG>::WARPS * 32)
assemble_kernel(const AssembleArgs a) {
  using PG = PassGeom<RIG>;
  constexpr int TPB = PG::TPB, BPW = PG::BPW, NCOL = PG::NCOL, BS = PG::BLK_STRIDE;
  extern __shared__ double smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk_id = blockIdx.x * PG::WARPS + warp;
  if (chunk_id >= a.n_chunks) return;  // whole warp leaves together
  double* rows = smem + warp * PG::WARP_SMEM;

  const Chunk ch = a.chunks[chunk_id];
  const int b = lane / TPB;
  const int t = lane - b * TPB;
  const int gi_id = tile_I(RIG, EPASS, t), gj_id = tile_J(RIG, EPASS, t);
  const bool lane_on = (lane < BPW * TPB) && (gi_id != G_NONE);
  const int gi = gi_id * 6, gj = gj_id * 6;
  const bool eval_lane = (lane < BPW * TPB) && (t < 4);

  double acc[36];
#pragma unroll
  for (int i = 0; i < 36; ++i) acc[i] = 0.0;

  const double* sh = a.shared + ch.cam * PG::SP;
  const double* xx = RIG ? a.ext_x + ch.cam * POSEX : nullptr;
  double* blk = rows + b * BS;

  for (int it = 0; it < ch.count; it += BPW) {
    const bool valid = (it + b) < ch.count;
    const int64_t g = (int64_t)ch.start + it + b;
    // ---- phase 1: corner evaluation ---------------------------------------
    if (valid && eval_lane) {
      const int oth = a.oth[g];
      const int vi = OWN_IS_VIEW ? ch.own : oth;
      const int mi = OWN_IS_VIEW ? oth : ch.own;
      BlockGeom<RIG> geo;
      block_geometry<RIG>(a.view_x + (size_t)vi * POSEX, a.marker_x + (size_t)mi * POSEX, xx, geo);
      double ox, oy;
      corner_xy(t, 0.5 * a.sizes[mi], ox, oy);
      const double2 px = *reinterpret_cast<const double2*>(a.pix + g * 8 + 2 * t);
      CornerRows<RIG> c;
      eval_corner<RIG, true>(geo, sh, ox, oy, px.x, px.y, c);
      if (!(c.depth > 0.0) || !isfinite(c.r[0]) || !isfinite(c.r[1])) *a.fail_flag = 1;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        double2* row = reinterpret_cast<double2*>(blk + (2 * t + i) * NCOL);
        const double* jo = OWN_IS_VIEW ? c.jv[i] : c.jm[i];
        const double* jt = OWN_IS_VIEW ? c.jm[i] : c.jv[i];
        row[0] = make_double2(jo[0], jo[1]);
        row[1] = make_double2(jo[2], jo[3]);
        row[2] = make_double2(jo[4], jo[5]);
        row[3] = make_double2(jt[0], jt[1]);
        row[4] = make_double2(jt[2], jt[3]);
        row[5] = make_double2(jt[4], jt[5]);
        row[6] = make_double2(c.js[i][0], c.js[i][1]);
        row[7] = make_double2(c.js[i][2], c.js[i][3]);
        row[8] = make_double2(c.js[i][4], c.js[i][5]);
        row[9] = make_double2(c.js[i][6], c.js[i][7]);
        row[10] = make_double2(c.js[i][8], c.r[i]);
        row[11] = make_double2(0.0, 0.0);
        if (RIG) {
          row[12] = make_double2(c.jx[i][0], c.jx[i][1]);
          row[13] = make_double2(c.jx[i][2], c.jx[i][3]);
          row[14] = make_double2(c.jx[i][4], c.jx[i][5]);
        }
      }
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
  for (int k = 0; k < 36; ++k) rows[k * 32 + lane] = acc[k];
  __syncwarp();
  double* out = a.partials + (size_t)chunk_id * (TPB * 36);
  for (int o = lane; o < TPB * 36; o += 32) {
    const int tt = o / 36, k = o - tt * 36;
    double s = 0.0;
#pragma unroll
    for (int bb = 0; bb < BPW; ++bb) s += rows[k * 32 + bb * TPB + tt];
    out[o] = s;
  }
}

template <bool RIG, bool EPASS, bool OWN_IS_VIEW>
static void launch_assemble_t(const AssembleArgs& a, cudaStream_t s) {
  using PG = PassGeom<RIG>;
  if (a.n_chunks == 0) return;
  const size_t smem = PG::WARPS * PG::WARP_SMEM * sizeof(double);
  auto k = assemble_kernel<RIG, EPASS, OWN_IS_VIEW>;
  static bool attr_set = false;
  if (!attr_set) {
    RCC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int grid = ceil_div(a.n_chunks, PG::WARPS);
  k<<<grid, PG::WARPS * 32, smem, s>>>(a);
  RCC_CUDA(cudaGetLastError());
}

void launch_assemble(bool rig, bool epass, bool own_is_view, const AssembleArgs& a, cudaStream_t s) {
  const int sel = (rig ? 4 : 0) | (epass ? 2 : 0) | (own_is_view ? 1 : 0);
  switch (sel) {
    case 0: launch_assemble_t<false, false, false>(a, s); break;
    case 1: launch_assemble_t<false, false, true>(a, s); break;
    case 2: launch_assemble_t<false, true, false>(a, s); break;
    case 3: launch_assemble_t<false, true, true>(a, s); break;
    case 4: launch_assemble_t<true, false, false>(a, s); break;
    case 5: launch_assemble_t<true, false, true>(a, s); break;
    case 6: launch_assemble_t<true, true, false>(a, s); break;
    default: launch_assemble_t<true, true, true>(a, s); break;
  }
}

// ---------------------------------------------------------------------------
// finalize: one warp per own block sums its chunk partials in chunk order.
// ---------------------------------------------------------------------------
template <bool RIG, bool EPASS>
__global__ void __launch_bounds__(128) finalize_side_kernel(const FinalizeSideArgs a) {
  using PG = PassGeom<RIG>;
  constexpr int TPB = PG::TPB, SP = PG::SP;
  constexpr int T_S1 = EPASS ? 2 : 1, T_S2 = T_S1 + 1, T_X = T_S1 + 2;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= a.n_own) return;
  double* hos = a.Hos + (size_t)i * 6 * a.n_shared;
  for (int k = lane; k < 6 * a.n_shared; k += 32) hos[k] = 0.0;
  __syncwarp();
  double oo[2] = {0.0, 0.0}, gacc[2] = {0.0, 0.0};
  for (int c = a.chunk_ptr[i]; c < a.chunk_ptr[i + 1]; ++c) {
    const double* P = a.partials + (size_t)c * (TPB * 36);
    const int cam = a.chunks[c].cam;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = lane + 32 * h;
      if (k < 36) {
        const int r = k / 6, col = k - 6 * r;
        oo[h] += P[k];
        double* hrow = hos + r * a.n_shared + cam * SP;
        hrow[col] += P[T_S1 * 36 + k];
        if (col < 3) hrow[6 + col] += P[T_S2 * 36 + k];
        else if (col == 3) gacc[h] += P[T_S2 * 36 + k];
        if (RIG) hrow[9 + col] += P[T_X * 36 + k];
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int k = lane + 32 * h;
    if (k < 36) {
      a.Hoo[(size_t)i * 36 + k] = oo[h];
      const int r = k / 6, col = k - 6 * r;
      if (col == 3) a.go[(size_t)i * 6 + r] = gacc[h];
    }
  }
}

void launch_finalize_side(bool rig, bool epass, const FinalizeSideArgs& a, cudaStream_t s) {
  if (a.n_own == 0) return;
  const int grid = ceil_div((int64_t)a.n_own * 32, 128);
  if (rig) {
    if (epass) finalize_side_kernel<true, true><<<grid, 128, 0, s>>>(a);
    else finalize_side_kernel<true, false><<<grid, 128, 0, s>>>(a);
  } else {
    if (epass) finalize_side_kernel<false, true><<<grid, 128, 0, s>>>(a);
    else finalize_side_kernel<false, false><<<grid, 128, 0, s>>>(a);
  }
  RCC_CUDA(cudaGetLastError());
}

// CTA-wide fixed-order reduction of one tile over a chunk list -> out36 (smem)
__device__ void reduce_tile(const double* __restrict__ partials, int tpb, int tile, const int32_t* __restrict__ list,
                            int n_list, double* red /*[7*36]*/, double* out36) {
  const int tid = threadIdx.x;
  const int slice = tid / 36, k = tid - 36 * slice;
  if (slice < 7) {
    double s = 0.0;
    for (int q = slice; q < n_list; q += 7) s += partials[(size_t)list[q] * (tpb * 36) + tile * 36 + k];
    red[slice * 36 + k] = s;
  }
  __syncthreads();
  if (tid < 36) {
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < 7; ++q) s += red[q * 36 + tid];
    out36[tid] = s;
  }
  __syncthreads();
}

template <bool RIG>
__global__ void __launch_bounds__(256) finalize_shared_kernel(const FinalizeSharedArgs a) {
  using PG = PassGeom<RIG>;
  constexpr int TPB = PG::TPB, SP = PG::SP;
  __shared__ double red[7 * 36];
  __shared__ double tile[36];
  const int cam = blockIdx.x;
  const int tid = threadIdx.x;
  const int ns = a.n_shared;
  double* H = a.Hss;
  const int base = cam * SP;
  for (int k = tid; k < SP * ns; k += blockDim.x) H[(size_t)base * ns + k] = 0.0;
  __syncthreads();
  const int32_t* le = a.cam_chunks_e + a.cam_ptr_e[cam];
  const int ne = a.cam_ptr_e[cam + 1] - a.cam_ptr_e[cam];
  const int32_t* lf = a.cam_chunks_f + a.cam_ptr_f[cam];
  const int nf = a.cam_ptr_f[cam + 1] - a.cam_ptr_f[cam];
  const int r = tid / 6, q = tid - 6 * r;  // valid for tid < 36

  // S1 x S1  (E pass)
  reduce_tile(a.part_e, TPB, RIG ? 5 : 4, le, ne, red, tile);
  if (tid < 36) H[(size_t)(base + r) * ns + base + q] = tile[tid];
  __syncthreads();
  // S1 x S2  (F pass): cols 0..2 -> p1 p2 k3, col 3 -> gradient of S1
  reduce_tile(a.part_f, TPB, RIG ? 4 : 3, lf, nf, red, tile);
  if (tid < 36) {
    if (q < 3) {
      H[(size_t)(base + r) * ns + base + 6 + q] = tile[tid];
      H[(size_t)(base + 6 + q) * ns + base + r] = tile[tid];
    } else if (q == 3) {
      a.gs[base + r] = tile[tid];
    }
  }
  __syncthreads();
  // S2 x S2  (F pass)
  reduce_tile(a.part_f, TPB, RIG ? 5 : 4, lf, nf, red, tile);
  if (tid < 36) {
    if (r < 3 && q < 3) H[(size_t)(base + 6 + r) * ns + base + 6 + q] = tile[tid];
    else if (r < 3 && q == 3) a.gs[base + 6 + r] = tile[tid];
    else if (r == 3 && q == 3) a.cost2_cam[cam] = tile[tid];
  }
  __syncthreads();
  if (RIG) {
    // S1 x X (E pass)
    reduce_tile(a.part_e, TPB, 6, le, ne, red, tile);
    if (tid < 36) {
      H[(size_t)(base + r) * ns + base + 9 + q] = tile[tid];
      H[(size_t)(base + 9 + q) * ns + base + r] = tile[tid];
    }
    __syncthreads();
    // X x X (E pass)
    reduce_tile(a.part_e, TPB, 7, le, ne, red, tile);
    if (tid < 36) H[(size_t)(base + 9 + r) * ns + base + 9 + q] = tile[tid];
    __syncthreads();
    // S2 x X (F pass): rows 0..2 -> p1 p2 k3, row 3 -> gradient of X
    reduce_tile(a.part_f, TPB, 6, lf, nf, red, tile);
    if (tid < 36) {
      if (r < 3) {
        H[(size_t)(base + 6 + r) * ns + base + 9 + q] = tile[tid];
        H[(size_t)(base + 9 + q) * ns + base + 6 + r] = tile[tid];
      } else if (r == 3) {
        a.gs[base + 9 + q] = tile[tid];
      }
    }
  }
}

void launch_finalize_shared(bool rig, const FinalizeSharedArgs& a, cudaStream_t s) {
  if (rig) finalize_shared_kernel<true><<<a.n_cam, 256, 0, s>>>(a);
  else finalize_shared_kernel<false><<<a.n_cam, 256, 0, s>>>(a);
  RCC_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
__global__ void expand_poses_kernel(const double* __restrict__ views, int n_views, double* __restrict__ view_x,
                                    const double* __restrict__ markers, int n_markers, double* __restrict__ marker_x,
                                    const double* __restrict__ shared, int n_cam, int sp, double* __restrict__ ext_x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_views) {
    expand_pose(views + (size_t)i * 6, view_x + (size_t)i * POSEX);
  } else if (i < n_views + n_markers) {
    const int j = i - n_views;
    expand_pose(markers + (size_t)j * 6, marker_x + (size_t)j * POSEX);
  } else if (i < n_views + n_markers + n_cam && sp == 15) {
    const int j = i - n_views - n_markers;
    expand_pose(shared + (size_t)j * sp + 9, ext_x + (size_t)j * POSEX);
  }
}

void launch_expand_poses(const double* views, int n_views, double* view_x, const double* markers, int n_markers,
                         double* marker_x, const double* shared, int n_cam, int sp, double* ext_x, cudaStream_t s) {
  const int n = n_views + n_markers + n_cam;
  expand_poses_kernel<<<ceil_div(n, 128), 128, 0, s>>>(views, n_views, view_x, markers, n_markers, marker_x, shared,
                                                       n_cam, sp, ext_x);
  RCC_CUDA(cudaGetLastError());
}

}  // namespace rcc
