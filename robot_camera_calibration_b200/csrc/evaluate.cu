// K1 -- materialised cost-functor evaluation (residuals + per-parameter-block
// Jacobians in the Ceres CostFunction::Evaluate layout), and K6's cost-only
// evaluation used for the LM gain ratio.
//
// One thread per tag corner (4 threads = one observation block); each thread
// writes its two residual rows of every Jacobian block as 16-byte vector
// stores.  The kernel is HBM-write-bound: 1 408 B out per 72 B in.  Rows are staged in shared
// memory in the Ceres layout and leave the SM as TMA bulk copies (cp.async.bulk, one per
// (tag, Jacobian block) record).
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

// TMA bulk copy shared::cta -> global (16-byte aligned, size a multiple of 16); SASS: UBLKCP
__device__ __forceinline__ void bulk_store(double* gdst, const double* ssrc, int bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
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
    // expanded pose records as 16-byte loads (the 4 corner lanes of a tag read the same lines)
    double vx[POSEX], mx[POSEX], xx[POSEX];
    {
      const double2* pv = reinterpret_cast<const double2*>(a.view_x + (size_t)vi * POSEX);
      const double2* pm = reinterpret_cast<const double2*>(a.marker_x + (size_t)mi * POSEX);
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        const double2 u = pv[k], w = pm[k];
        vx[2 * k] = u.x; vx[2 * k + 1] = u.y;
        mx[2 * k] = w.x; mx[2 * k + 1] = w.y;
      }
      if (RIG) {
        const double2* pe = reinterpret_cast<const double2*>(a.ext_x + (size_t)cam * POSEX);
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          const double2 u = pe[k];
          xx[2 * k] = u.x; xx[2 * k + 1] = u.y;
        }
      }
    }
    BlockGeom<RIG> geo;
    block_geometry<RIG>(vx, mx, RIG ? xx : nullptr, geo);
    double ox, oy;
    corner_xy(t, mx[PX_HS], ox, oy);
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
    // hand the copy-out to the TMA engine: one bulk copy (shared -> global) per (block, array)
    // record, 64..384 contiguous bytes each, issued by one lane per record
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // staged rows -> visible to the async proxy
    __syncwarp();
    const int64_t g0 = ((int64_t)blockIdx.x * EVAL_THREADS + warp * 32) >> 2;
    const int nblk = (int)max((int64_t)0, min((int64_t)8, a.n - g0));
    constexpr int NARR = RIG ? 6 : 5;
    for (int idx = lane; idx < nblk * NARR; idx += 32) {
      const int q = idx / NARR, arr = idx - q * NARR;
      double* dst = nullptr;
      int off = 0, k = 0;
      switch (arr) {
        case 0: dst = a.residuals; off = ER::RES; k = 8; break;
        case 1: dst = a.jac_intr; off = ER::JI; k = 32; break;
        case 2: dst = a.jac_dist; off = ER::JD; k = 40; break;
        case 3: dst = a.jac_view; off = ER::JV; k = 48; break;
        case 4: dst = a.jac_marker; off = ER::JM; k = 48; break;
        default: dst = a.jac_ext; off = ER::JX; k = 48; break;
      }
      if (dst) bulk_store(dst + wpos[q] * k, wrec + q * ER::SIZE + off, k * (int)sizeof(double));
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the reads
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
