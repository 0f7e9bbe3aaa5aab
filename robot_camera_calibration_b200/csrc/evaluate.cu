// K1 -- materialised cost-functor evaluation (residuals + per-parameter-block
// Jacobians in the Ceres CostFunction::Evaluate layout), and K6's cost-only
// evaluation used for the LM gain ratio.
//
// One thread per tag corner (4 threads = one observation block); each thread
// writes its two residual rows of every Jacobian block as 16-byte vector
// stores.  The kernel is HBM-write-bound: 1 408 B out per 72 B in.
#include "common.cuh"
#include "kernels.h"
#include "model.cuh"

namespace rcc {

constexpr int EVAL_THREADS = 256;

int eval_grid(int64_t n) { return ceil_div(n * 4, EVAL_THREADS); }

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  }
  return s;  // valid on thread 0
}

// per-block staging record (doubles): res[8] | intr[8][4] | dist[8][5] | view[8][6] | marker[8][6] | ext[8][6]
template <bool RIG>
struct EvalRec {
  static constexpr int RES = 0, JI = 8, JD = 40, JV = 80, JM = 128, JX = 176;
  static constexpr int SIZE = RIG ? 226 : 178;   // +2: consecutive records shift by 16 B across the smem banks
};

// cooperative, coalesced copy-out of one output array for the 8 blocks of a warp:
// K doubles per block, contiguous in global memory per block (caller order via orig)
template <int K, int REC>
__device__ __forceinline__ void copy_out(double* __restrict__ dst, const double* wrec, int off, const int64_t* o8,
                                         int nblk, int lane) {
  constexpr int PIECES = K / 2;  // 16-byte pieces per block
  for (int idx = lane; idx < nblk * PIECES; idx += 32) {
    const int q = idx / PIECES, pc = idx - q * PIECES;
    const double2 v = *reinterpret_cast<const double2*>(wrec + q * REC + off + 2 * pc);
    *reinterpret_cast<double2*>(dst + o8[q] * K + 2 * pc) = v;
  }
}

template <bool RIG, bool WANT_J>
__global__ void __launch_bounds__(EVAL_THREADS) evaluate_kernel(const EvalArgs a) {
  using ER = EvalRec<RIG>;
  __shared__ double red[EVAL_THREADS / 32];
  extern __shared__ __align__(16) double stage[];   // WANT_J: [warps][8 blocks][ER::SIZE]; [warps][8] caller positions
  const int64_t tid = (int64_t)blockIdx.x * EVAL_THREADS + threadIdx.x;
  const int64_t g = tid >> 2;
  const int t = (int)(tid & 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double r2 = 0.0;
  double* wrec = WANT_J ? stage + (size_t)warp * 8 * ER::SIZE : nullptr;
  int64_t* wpos = WANT_J ? reinterpret_cast<int64_t*>(stage + (size_t)(EVAL_THREADS / 32) * 8 * ER::SIZE) + warp * 8 : nullptr;
  if (g < a.n) {
    const int vi = a.view_idx[g], mi = a.marker_idx[g], cam = a.cam[g];
    constexpr int SP = RIG ? 15 : 9;
    BlockGeom<RIG> geo;
    block_geometry<RIG>(a.view_x + (size_t)vi * POSEX, a.marker_x + (size_t)mi * POSEX,
                        RIG ? a.ext_x + (size_t)cam * POSEX : nullptr, geo);
    double ox, oy;
    corner_xy(t, 0.5 * a.sizes[mi], ox, oy);
    const double2 px = *reinterpret_cast<const double2*>(a.pix + g * 8 + 2 * t);
    CornerRows<RIG> c;
    eval_corner<RIG, WANT_J>(geo, a.shared + (size_t)cam * SP, ox, oy, px.x, px.y, c);
    if (!(c.depth > 0.0) || !isfinite(c.r[0]) || !isfinite(c.r[1])) *a.fail_flag = 1;
    r2 = c.r[0] * c.r[0] + c.r[1] * c.r[1];
    const int64_t o = a.orig ? (int64_t)a.orig[g] : g;
    if (!WANT_J) {
      if (a.residuals) *reinterpret_cast<double2*>(a.residuals + o * 8 + 2 * t) = make_double2(c.r[0], c.r[1]);
    } else {
      // stage the corner's two rows of every block in Ceres layout
      double* rec = wrec + (lane >> 2) * ER::SIZE;
      if (t == 0) wpos[lane >> 2] = o;
      *reinterpret_cast<double2*>(rec + ER::RES + 2 * t) = make_double2(c.r[0], c.r[1]);
      double2* d = reinterpret_cast<double2*>(rec + ER::JI + 8 * t);
      d[0] = make_double2(c.js[0][0], c.js[0][1]);
      d[1] = make_double2(c.js[0][2], c.js[0][3]);
      d[2] = make_double2(c.js[1][0], c.js[1][1]);
      d[3] = make_double2(c.js[1][2], c.js[1][3]);
      d = reinterpret_cast<double2*>(rec + ER::JD + 10 * t);
      d[0] = make_double2(c.js[0][4], c.js[0][5]);
      d[1] = make_double2(c.js[0][6], c.js[0][7]);
      d[2] = make_double2(c.js[0][8], c.js[1][4]);
      d[3] = make_double2(c.js[1][5], c.js[1][6]);
      d[4] = make_double2(c.js[1][7], c.js[1][8]);
      d = reinterpret_cast<double2*>(rec + ER::JV + 12 * t);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) d[3 * i + j] = make_double2(c.jv[i][2 * j], c.jv[i][2 * j + 1]);
      d = reinterpret_cast<double2*>(rec + ER::JM + 12 * t);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) d[3 * i + j] = make_double2(c.jm[i][2 * j], c.jm[i][2 * j + 1]);
      if (RIG) {
        d = reinterpret_cast<double2*>(rec + ER::JX + 12 * t);
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) d[3 * i + j] = make_double2(c.jx[i][2 * j], c.jx[i][2 * j + 1]);
      }
    }
  }
  if (WANT_J) {
    __syncwarp();
    // blocks of this warp: g0 .. g0+7 (clipped)
    const int64_t g0 = ((int64_t)blockIdx.x * EVAL_THREADS + warp * 32) >> 2;
    const int nblk = (int)max((int64_t)0, min((int64_t)8, a.n - g0));
    if (nblk > 0) {
      if (a.residuals) copy_out<8, ER::SIZE>(a.residuals, wrec, ER::RES, wpos, nblk, lane);
      if (a.jac_intr) copy_out<32, ER::SIZE>(a.jac_intr, wrec, ER::JI, wpos, nblk, lane);
      if (a.jac_dist) copy_out<40, ER::SIZE>(a.jac_dist, wrec, ER::JD, wpos, nblk, lane);
      if (a.jac_view) copy_out<48, ER::SIZE>(a.jac_view, wrec, ER::JV, wpos, nblk, lane);
      if (a.jac_marker) copy_out<48, ER::SIZE>(a.jac_marker, wrec, ER::JM, wpos, nblk, lane);
      if (RIG && a.jac_ext) copy_out<48, ER::SIZE>(a.jac_ext, wrec, ER::JX, wpos, nblk, lane);
    }
  }
  if (a.loss != 0) {
    // the 4 corner threads of a tag are adjacent lanes: rho(sum of the 8 squared residuals) / 4 each
    double sb = r2 + __shfl_xor_sync(0xffffffffu, r2, 1);
    sb += __shfl_xor_sync(0xffffffffu, sb, 2);
    double rho1;
    r2 = (g < a.n) ? 0.25 * robust_rho(a.loss, a.loss_a2, sb, rho1) : 0.0;
  }
  const double s = block_sum(r2, red);
  if (threadIdx.x == 0 && a.cost2_partials) a.cost2_partials[blockIdx.x] = s;
}

template <bool RIG, bool WANT_J>
static void launch_evaluate_t(const EvalArgs& a, cudaStream_t s) {
  const int grid = eval_grid(a.n);
  size_t smem = 0;
  if (WANT_J) smem = (size_t)(EVAL_THREADS / 32) * 8 * (EvalRec<RIG>::SIZE * sizeof(double) + sizeof(int64_t));
  auto k = evaluate_kernel<RIG, WANT_J>;
  static bool attr = false;
  if (!attr && smem > 48 * 1024) {
    RCC_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  k<<<grid, EVAL_THREADS, smem, s>>>(a);
  RCC_CUDA(cudaGetLastError());
}

void launch_evaluate(bool rig, bool want_jac, const EvalArgs& a, cudaStream_t s) {
  if (a.n == 0) return;
  if (rig) {
    if (want_jac) launch_evaluate_t<true, true>(a, s);
    else launch_evaluate_t<true, false>(a, s);
  } else {
    if (want_jac) launch_evaluate_t<false, true>(a, s);
    else launch_evaluate_t<false, false>(a, s);
  }
}

void launch_cost(bool rig, const EvalArgs& a, cudaStream_t s) {
  EvalArgs b = a;
  b.residuals = nullptr;
  b.jac_intr = b.jac_dist = b.jac_view = b.jac_marker = b.jac_ext = nullptr;
  launch_evaluate(rig, false, b, s);
}

// deterministic single-CTA sum
__global__ void __launch_bounds__(1024) sum_kernel(const double* __restrict__ in, int64_t n, double* __restrict__ out,
                                                   double scale) {
  __shared__ double red[1024];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += in[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = red[0] * scale;
}

void launch_sum(const double* in, int64_t n, double* out, double scale, cudaStream_t s) {
  sum_kernel<<<1, 1024, 0, s>>>(in, n, out, scale);
  RCC_CUDA(cudaGetLastError());
}

}  // namespace rcc
