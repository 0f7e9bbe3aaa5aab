// "Ceres-equivalent" CPU restatement of the bundle-adjustment hot path.
// TEST / BASELINE INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library.  The
// product (robot_camera_calibration_b200/) never does.
//
// PARITY STATUS: unpinned at the Ceres boundary.  Ceres Solver is neither in
// /root/reference nor installed here; this file restates, from
// the published Ceres design, how a Ceres user of the reference's pipeline
// would evaluate this problem on the CPU:
//   * the cost functor is a template over the scalar type and is differentiated
//     with forward-mode dual numbers ("Jets"), as ceres::AutoDiffCostFunction does;
//     one residual block = one tag = 4 corners = 8 residuals
//     (corner_detections.cpp:34,51), parameter blocks intr[4] dist[5] view[6]
//     marker[6] (+ext[6] for a rig) with the conventions of camera_pose.cpp:38-39,
//     55-68 (K, dist), :88-98,111-121 (Rodrigues poses), :123-126 (corner order);
//   * CostFunction::Evaluate fills row-major per-parameter-block Jacobians;
//   * the block-sparse Jacobian is turned into J^T J / J^T r blocks, the
//     eliminated 6-dof blocks are removed by a block Schur complement into a
//     dense reduced matrix (per-cell locks, like Ceres's dense random-access
//     block matrix), all parallel over the eliminated blocks with OpenMP.
// The dense Cholesky of the reduced system is left to LAPACK via SciPy in
// oracle/cpu_baseline.py (what Ceres's DENSE_SCHUR + LAPACK backend would call).
//
// Because the derivative here comes from dual numbers and the GPU path uses
// hand-derived analytic Jacobians, this file is also an independent check of
// the GPU arithmetic (tests/test_cpu_restatement.py pins it to the numpy
// complex-step oracle and to OpenCV's projectPoints derivatives).
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <vector>

namespace {

// ---------------------------------------------------------------- dual numbers
template <int N>
struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  explicit Jet(double x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double x, int k) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1; }
};
template <int N> inline Jet<N> operator+(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a + y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a - y.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
template <int N> inline Jet<N> operator-(const Jet<N>& x) { Jet<N> r; r.a = -x.a; for (int i = 0; i < N; ++i) r.v[i] = -x.v[i]; return r; }
template <int N> inline Jet<N> operator*(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; r.a = x.a * y.a; for (int i = 0; i < N; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
template <int N> inline Jet<N> operator/(const Jet<N>& x, const Jet<N>& y) { Jet<N> r; const double iy = 1.0 / y.a; r.a = x.a * iy; for (int i = 0; i < N; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * iy; return r; }
template <int N> inline Jet<N> operator+(const Jet<N>& x, double s) { Jet<N> r = x; r.a += s; return r; }
template <int N> inline Jet<N> operator+(double s, const Jet<N>& x) { return x + s; }
template <int N> inline Jet<N> operator-(const Jet<N>& x, double s) { Jet<N> r = x; r.a -= s; return r; }
template <int N> inline Jet<N> operator*(const Jet<N>& x, double s) { Jet<N> r; r.a = x.a * s; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * s; return r; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& x) { return x * s; }
template <int N> inline Jet<N> sqrt(const Jet<N>& x) { Jet<N> r; r.a = ::sqrt(x.a); const double d = 0.5 / r.a; for (int i = 0; i < N; ++i) r.v[i] = x.v[i] * d; return r; }
template <int N> inline Jet<N> sin(const Jet<N>& x) { Jet<N> r; r.a = ::sin(x.a); const double c = ::cos(x.a); for (int i = 0; i < N; ++i) r.v[i] = c * x.v[i]; return r; }
template <int N> inline Jet<N> cos(const Jet<N>& x) { Jet<N> r; r.a = ::cos(x.a); const double s = -::sin(x.a); for (int i = 0; i < N; ++i) r.v[i] = s * x.v[i]; return r; }
inline double sqrt(double x) { return ::sqrt(x); }
inline double sin(double x) { return ::sin(x); }
inline double cos(double x) { return ::cos(x); }
inline double value(double x) { return x; }
template <int N> inline double value(const Jet<N>& x) { return x.a; }

// ---------------------------------------------------------------- cost functor
// rotate pt by the angle-axis vector aa (Rodrigues formula; first-order branch
// near zero so the derivative at aa = 0 is exact: d(R p)/d aa = -[p]x)
template <typename T>
inline void angle_axis_rotate(const T aa[3], const T pt[3], T out[3]) {
  const T t2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (value(t2) > 1e-20) {
    const T t = sqrt(t2);
    const T c = cos(t), s = sin(t);
    const T it = T(1.0) / t;
    const T w[3] = {aa[0] * it, aa[1] * it, aa[2] * it};
    const T wxp[3] = {w[1] * pt[2] - w[2] * pt[1], w[2] * pt[0] - w[0] * pt[2], w[0] * pt[1] - w[1] * pt[0]};
    const T k = (w[0] * pt[0] + w[1] * pt[1] + w[2] * pt[2]) * (T(1.0) - c);
    for (int i = 0; i < 3; ++i) out[i] = pt[i] * c + wxp[i] * s + w[i] * k;
  } else {
    out[0] = pt[0] + (aa[1] * pt[2] - aa[2] * pt[1]);
    out[1] = pt[1] + (aa[2] * pt[0] - aa[0] * pt[2]);
    out[2] = pt[2] + (aa[0] * pt[1] - aa[1] * pt[0]);
  }
}

// residuals of one tag (4 corners) -- the functor a Ceres user would template
template <typename T, bool RIG>
inline bool reprojection_functor(const T* intr, const T* dist, const T* view, const T* marker, const T* ext,
                                 double size, const double* pix, T* res) {
  const double h = 0.5 * size;
  const double ox[4] = {-h, h, h, -h}, oy[4] = {-h, -h, h, h};   // bl br tr tl
  const T nview[3] = {-view[0], -view[1], -view[2]};
  bool ok = true;
  for (int k = 0; k < 4; ++k) {
    const T o[3] = {T(ox[k]), T(oy[k]), T(0.0)};
    T pw[3];
    angle_axis_rotate(marker, o, pw);                              // world_T_target
    const T q[3] = {pw[0] + marker[3] - view[3], pw[1] + marker[4] - view[4], pw[2] + marker[5] - view[5]};
    T pc[3];
    angle_axis_rotate(nview, q, pc);                               // inverse of world_T_camera|body
    if (RIG) {
      const T next[3] = {-ext[0], -ext[1], -ext[2]};
      const T qb[3] = {pc[0] - ext[3], pc[1] - ext[4], pc[2] - ext[5]};
      angle_axis_rotate(next, qb, pc);                             // inverse of body_T_cam
    }
    if (!(value(pc[2]) > 0.0)) ok = false;
    const T x = pc[0] / pc[2], y = pc[1] / pc[2];
    const T r2 = x * x + y * y;
    const T rad = T(1.0) + r2 * (dist[0] + r2 * (dist[1] + r2 * dist[4]));
    const T xd = x * rad + T(2.0) * dist[2] * x * y + dist[3] * (r2 + T(2.0) * x * x);
    const T yd = y * rad + dist[2] * (r2 + T(2.0) * y * y) + T(2.0) * dist[3] * x * y;
    res[2 * k] = intr[0] * xd + intr[2] - pix[2 * k];
    res[2 * k + 1] = intr[1] * yd + intr[3] - pix[2 * k + 1];
  }
  return ok;
}

struct Problem {
  int rig, n_views, n_markers, n_cam, elim_view, sp, n_shared, n_e, n_f;
  int64_t n;
  std::vector<int32_t> vi, mi, ci;
  std::vector<double> pix, views, markers, sizes, intr, dist, ext;
  // grouping
  std::vector<int32_t> e_ptr, e_list, f_ptr, f_list;
  // evaluation products (Ceres layout: per residual block, row-major per parameter block)
  std::vector<double> res, ji, jd, jv, jm, jx;
  // normal equation blocks
  std::vector<double> Hee, ge, Hes, Hff, gf, Hfs, Hss, gs, W;
  double cost;
  int fail;
};

template <bool RIG>
void evaluate_all(Problem& P, bool want_j) {
  constexpr int N = RIG ? 27 : 21;
  typedef Jet<N> J;
  const int64_t n = P.n;
  int fail = 0;
#pragma omp parallel for schedule(static) reduction(| : fail)
  for (int64_t b = 0; b < n; ++b) {
    const int v = P.vi[b], m = P.mi[b], c = P.ci[b];
    const double* pv = &P.views[(size_t)v * 6];
    const double* pm = &P.markers[(size_t)m * 6];
    const double* pi = &P.intr[(size_t)c * 4];
    const double* pd = &P.dist[(size_t)c * 5];
    const double* pe = &P.ext[(size_t)c * 6];
    if (!want_j) {
      double r[8];
      if (!reprojection_functor<double, RIG>(pi, pd, pv, pm, pe, P.sizes[m], &P.pix[(size_t)b * 8], r)) fail |= 1;
      memcpy(&P.res[(size_t)b * 8], r, 64);
      continue;
    }
    J ji[4], jd[5], jv[6], jm[6], je[6], r[8];
    int k = 0;
    for (int i = 0; i < 4; ++i) ji[i] = J(pi[i], k++);
    for (int i = 0; i < 5; ++i) jd[i] = J(pd[i], k++);
    for (int i = 0; i < 6; ++i) jv[i] = J(pv[i], k++);
    for (int i = 0; i < 6; ++i) jm[i] = J(pm[i], k++);
    if (RIG) for (int i = 0; i < 6; ++i) je[i] = J(pe[i], k++);
    if (!reprojection_functor<J, RIG>(ji, jd, jv, jm, je, P.sizes[m], &P.pix[(size_t)b * 8], r)) fail |= 1;
    for (int row = 0; row < 8; ++row) {
      P.res[(size_t)b * 8 + row] = r[row].a;
      if (!std::isfinite(r[row].a)) fail |= 1;
      for (int i = 0; i < 4; ++i) P.ji[(size_t)b * 32 + row * 4 + i] = r[row].v[i];
      for (int i = 0; i < 5; ++i) P.jd[(size_t)b * 40 + row * 5 + i] = r[row].v[4 + i];
      for (int i = 0; i < 6; ++i) P.jv[(size_t)b * 48 + row * 6 + i] = r[row].v[9 + i];
      for (int i = 0; i < 6; ++i) P.jm[(size_t)b * 48 + row * 6 + i] = r[row].v[15 + i];
      if (RIG) for (int i = 0; i < 6; ++i) P.jx[(size_t)b * 48 + row * 6 + i] = r[row].v[21 + i];
    }
  }
  P.fail = fail;
}

// A^T B accumulate: A 8 x ka, B 8 x kb (row-major) -> C ka x ldc
inline void atb(const double* A, int ka, const double* B, int kb, double* C, int ldc) {
  for (int r = 0; r < 8; ++r)
    for (int i = 0; i < ka; ++i) {
      const double a = A[r * ka + i];
      for (int j = 0; j < kb; ++j) C[i * ldc + j] += a * B[r * kb + j];
    }
}
inline void atv(const double* A, int ka, const double* r8, double* g) {
  for (int r = 0; r < 8; ++r)
    for (int i = 0; i < ka; ++i) g[i] += A[r * ka + i] * r8[r];
}

void group(const std::vector<int32_t>& key, int nk, std::vector<int32_t>& ptr, std::vector<int32_t>& list) {
  ptr.assign((size_t)nk + 1, 0);
  for (size_t i = 0; i < key.size(); ++i) ptr[key[i] + 1]++;
  for (int i = 0; i < nk; ++i) ptr[i + 1] += ptr[i];
  list.resize(key.size());
  std::vector<int32_t> cur(ptr.begin(), ptr.end() - 1);
  for (size_t i = 0; i < key.size(); ++i) list[cur[key[i]]++] = (int32_t)i;
}

// shared-parameter Jacobian of block b as one 8 x sp row-major matrix
inline void shared_jac(const Problem& P, int64_t b, double* Js) {
  const int sp = P.sp;
  for (int r = 0; r < 8; ++r) {
    for (int i = 0; i < 4; ++i) Js[r * sp + i] = P.ji[(size_t)b * 32 + r * 4 + i];
    for (int i = 0; i < 5; ++i) Js[r * sp + 4 + i] = P.jd[(size_t)b * 40 + r * 5 + i];
    if (P.rig) for (int i = 0; i < 6; ++i) Js[r * sp + 9 + i] = P.jx[(size_t)b * 48 + r * 6 + i];
  }
}

void normal_blocks(Problem& P) {
  const int ns = P.n_shared, sp = P.sp;
  std::fill(P.Hee.begin(), P.Hee.end(), 0.0); std::fill(P.ge.begin(), P.ge.end(), 0.0);
  std::fill(P.Hes.begin(), P.Hes.end(), 0.0); std::fill(P.Hff.begin(), P.Hff.end(), 0.0);
  std::fill(P.gf.begin(), P.gf.end(), 0.0); std::fill(P.Hfs.begin(), P.Hfs.end(), 0.0);
  const std::vector<double>& Je = P.elim_view ? P.jv : P.jm;
  const std::vector<double>& Jf = P.elim_view ? P.jm : P.jv;
#pragma omp parallel for schedule(dynamic, 16)
  for (int e = 0; e < P.n_e; ++e) {
    double Js[8 * 15];
    for (int q = P.e_ptr[e]; q < P.e_ptr[e + 1]; ++q) {
      const int64_t b = P.e_list[q];
      const double* A = &Je[(size_t)b * 48];
      atb(A, 6, A, 6, &P.Hee[(size_t)e * 36], 6);
      atv(A, 6, &P.res[(size_t)b * 8], &P.ge[(size_t)e * 6]);
      double* w = &P.W[(size_t)b * 36];
      for (int i = 0; i < 36; ++i) w[i] = 0.0;
      atb(A, 6, &Jf[(size_t)b * 48], 6, w, 6);
      shared_jac(P, b, Js);
      atb(A, 6, Js, sp, &P.Hes[(size_t)e * 6 * ns + P.ci[b] * sp], ns);
    }
  }
#pragma omp parallel for schedule(dynamic, 16)
  for (int f = 0; f < P.n_f; ++f) {
    double Js[8 * 15];
    for (int q = P.f_ptr[f]; q < P.f_ptr[f + 1]; ++q) {
      const int64_t b = P.f_list[q];
      const double* A = &Jf[(size_t)b * 48];
      atb(A, 6, A, 6, &P.Hff[(size_t)f * 36], 6);
      atv(A, 6, &P.res[(size_t)b * 8], &P.gf[(size_t)f * 6]);
      shared_jac(P, b, Js);
      atb(A, 6, Js, sp, &P.Hfs[(size_t)f * 6 * ns + P.ci[b] * sp], ns);
    }
  }
  // shared x shared, gradient, cost: per-thread partials, fixed-order sum
  const int T = omp_get_max_threads();
  std::vector<double> part((size_t)T * (ns * ns + ns + 1), 0.0);
#pragma omp parallel
  {
    const int t = omp_get_thread_num();
    double* H = &part[(size_t)t * (ns * ns + ns + 1)];
    double* g = H + ns * ns;
    double* c = g + ns;
    double Js[8 * 15];
#pragma omp for schedule(static)
    for (int64_t b = 0; b < P.n; ++b) {
      shared_jac(P, b, Js);
      const int off = P.ci[b] * sp;
      atb(Js, sp, Js, sp, H + off * ns + off, ns);
      atv(Js, sp, &P.res[(size_t)b * 8], g + off);
      for (int r = 0; r < 8; ++r) *c += P.res[(size_t)b * 8 + r] * P.res[(size_t)b * 8 + r];
    }
  }
  std::fill(P.Hss.begin(), P.Hss.end(), 0.0);
  std::fill(P.gs.begin(), P.gs.end(), 0.0);
  double c2 = 0.0;
  for (int t = 0; t < T; ++t) {
    const double* H = &part[(size_t)t * (ns * ns + ns + 1)];
    for (int i = 0; i < ns * ns; ++i) P.Hss[i] += H[i];
    for (int i = 0; i < ns; ++i) P.gs[i] += H[ns * ns + i];
    c2 += H[ns * ns + ns];
  }
  P.cost = 0.5 * c2;
}

// 6x6 Cholesky solve helpers
inline bool chol6(const double* H, double* L) {
  for (int j = 0; j < 6; ++j) {
    double s = H[j * 6 + j];
    for (int k = 0; k < j; ++k) s -= L[j * 6 + k] * L[j * 6 + k];
    if (!(s > 0)) return false;
    L[j * 6 + j] = ::sqrt(s);
    for (int i = j + 1; i < 6; ++i) {
      double v = H[i * 6 + j];
      for (int k = 0; k < j; ++k) v -= L[i * 6 + k] * L[j * 6 + k];
      L[i * 6 + j] = v / L[j * 6 + j];
    }
    for (int i = 0; i < j; ++i) L[i * 6 + j] = 0.0;
  }
  return true;
}
// x := L^-1 x  (ncol columns, row-major 6 x ncol with leading dim ld)
inline void fwd6(const double* L, double* X, int ncol, int ld) {
  for (int i = 0; i < 6; ++i) {
    for (int k = 0; k < i; ++k) {
      const double l = L[i * 6 + k];
      for (int c = 0; c < ncol; ++c) X[i * ld + c] -= l * X[k * ld + c];
    }
    const double inv = 1.0 / L[i * 6 + i];
    for (int c = 0; c < ncol; ++c) X[i * ld + c] *= inv;
  }
}

}  // namespace

extern "C" {

int cpu_ba_num_threads() { return omp_get_max_threads(); }
// torchrun exports OMP_NUM_THREADS=1 and libgomp has read it long before this library is loaded: the CPU arm
// sets the team size at run time instead
void cpu_ba_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

void* cpu_ba_create(int rig, int n_views, int n_markers, int n_cam, int64_t n, int elim_view, const int32_t* vi,
                    const int32_t* mi, const int32_t* ci, const double* pix) {
  Problem* P = new Problem();
  P->rig = rig; P->n_views = n_views; P->n_markers = n_markers; P->n_cam = n_cam; P->n = n;
  P->elim_view = elim_view; P->sp = rig ? 15 : 9; P->n_shared = n_cam * P->sp;
  P->n_e = elim_view ? n_views : n_markers; P->n_f = elim_view ? n_markers : n_views;
  P->vi.assign(vi, vi + n); P->mi.assign(mi, mi + n);
  if (ci) P->ci.assign(ci, ci + n); else P->ci.assign((size_t)n, 0);
  P->pix.assign(pix, pix + n * 8);
  P->views.assign((size_t)n_views * 6, 0); P->markers.assign((size_t)n_markers * 6, 0);
  P->sizes.assign((size_t)n_markers, 0); P->intr.assign((size_t)n_cam * 4, 0); P->dist.assign((size_t)n_cam * 5, 0);
  P->ext.assign((size_t)n_cam * 6, 0);
  group(elim_view ? P->vi : P->mi, P->n_e, P->e_ptr, P->e_list);
  group(elim_view ? P->mi : P->vi, P->n_f, P->f_ptr, P->f_list);
  P->res.resize((size_t)n * 8); P->ji.resize((size_t)n * 32); P->jd.resize((size_t)n * 40);
  P->jv.resize((size_t)n * 48); P->jm.resize((size_t)n * 48); if (rig) P->jx.resize((size_t)n * 48);
  const int ns = P->n_shared;
  P->Hee.resize((size_t)P->n_e * 36); P->ge.resize((size_t)P->n_e * 6); P->Hes.resize((size_t)P->n_e * 6 * ns);
  P->Hff.resize((size_t)P->n_f * 36); P->gf.resize((size_t)P->n_f * 6); P->Hfs.resize((size_t)P->n_f * 6 * ns);
  P->Hss.resize((size_t)ns * ns); P->gs.resize(ns); P->W.resize((size_t)n * 36);
  P->cost = 0; P->fail = 0;
  return P;
}
void cpu_ba_destroy(void* h) { delete (Problem*)h; }

void cpu_ba_set_params(void* h, const double* views, const double* markers, const double* sizes, const double* intr,
                       const double* dist, const double* ext) {
  Problem& P = *(Problem*)h;
  if (views) P.views.assign(views, views + (size_t)P.n_views * 6);
  if (markers) P.markers.assign(markers, markers + (size_t)P.n_markers * 6);
  if (sizes) P.sizes.assign(sizes, sizes + (size_t)P.n_markers);
  if (intr) P.intr.assign(intr, intr + (size_t)P.n_cam * 4);
  if (dist) P.dist.assign(dist, dist + (size_t)P.n_cam * 5);
  if (ext) P.ext.assign(ext, ext + (size_t)P.n_cam * 6);
}

// CostFunction::Evaluate over every residual block.  returns 0 ok, 1 evaluation failure
int cpu_ba_evaluate(void* h, int want_j) {
  Problem& P = *(Problem*)h;
  if (P.rig) evaluate_all<true>(P, want_j != 0); else evaluate_all<false>(P, want_j != 0);
  return P.fail;
}
// Evaluate + J^T J / J^T r blocks
int cpu_ba_linearize(void* h, double* cost) {
  Problem& P = *(Problem*)h;
  if (P.rig) evaluate_all<true>(P, true); else evaluate_all<false>(P, true);
  normal_blocks(P);
  if (cost) *cost = P.cost;
  return P.fail;
}

void cpu_ba_get(void* h, double* res, double* ji, double* jd, double* jv, double* jm, double* jx, double* Hee,
                double* ge, double* Hes, double* Hff, double* gf, double* Hfs, double* Hss, double* gs, double* W) {
  Problem& P = *(Problem*)h;
  auto cp = [](double* d, const std::vector<double>& s) { if (d && !s.empty()) memcpy(d, s.data(), s.size() * 8); };
  cp(res, P.res); cp(ji, P.ji); cp(jd, P.jd); cp(jv, P.jv); cp(jm, P.jm); cp(jx, P.jx);
  cp(Hee, P.Hee); cp(ge, P.ge); cp(Hes, P.Hes); cp(Hff, P.Hff); cp(gf, P.gf); cp(Hfs, P.Hfs);
  cp(Hss, P.Hss); cp(gs, P.gs); cp(W, P.W);
}

// Block Schur complement into the dense reduced system (n_red x n_red row-major,
// upper triangle + mirrored by the caller) and rhs b.  Damping D^2 =
// clamp(diag)/radius on the eliminated blocks; e_const marks constant ones.
// Also returns what back-substitution needs: Linv-free form (we keep L and the
// transformed rows internally and expose back_substitute below).
struct SchurState {
  std::vector<double> L;    // n_e x 36
  std::vector<double> d2e;  // n_e x 6
};
static SchurState g_state;

int cpu_ba_schur(void* h, double radius, double min_diag, double max_diag, const uint8_t* e_const, double* S,
                 double* b) {
  Problem& P = *(Problem*)h;
  const int ns = P.n_shared, nf = P.n_f, n_red = 6 * nf + ns;
  g_state.L.assign((size_t)P.n_e * 36, 0.0);
  g_state.d2e.assign((size_t)P.n_e * 6, 0.0);
  // start from H_FF, g_F
  std::fill(S, S + (size_t)n_red * n_red, 0.0);
  for (int f = 0; f < nf; ++f) {
    for (int i = 0; i < 6; ++i) {
      for (int j = 0; j < 6; ++j) S[(size_t)(6 * f + i) * n_red + 6 * f + j] = P.Hff[(size_t)f * 36 + i * 6 + j];
      for (int s = 0; s < ns; ++s) S[(size_t)(6 * f + i) * n_red + 6 * nf + s] = P.Hfs[((size_t)f * 6 + i) * ns + s];
      b[6 * f + i] = P.gf[(size_t)f * 6 + i];
    }
  }
  for (int s = 0; s < ns; ++s) {
    for (int t = 0; t < ns; ++t) S[(size_t)(6 * nf + s) * n_red + 6 * nf + t] = P.Hss[(size_t)s * ns + t];
    b[6 * nf + s] = P.gs[s];
  }
  const int64_t ncell = (int64_t)(nf + 1) * (nf + 1);
  std::vector<std::atomic_flag> locks((size_t)ncell);
  for (auto& l : locks) l.clear();
  const std::vector<int32_t>& f_of = P.elim_view ? P.mi : P.vi;
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(| : bad)
  for (int e = 0; e < P.n_e; ++e) {
    if (e_const && e_const[e]) continue;
    const int m = P.e_ptr[e + 1] - P.e_ptr[e];
    double H[36], L[36];
    for (int i = 0; i < 36; ++i) H[i] = P.Hee[(size_t)e * 36 + i];
    for (int i = 0; i < 6; ++i) {
      const double d2 = std::min(std::max(H[i * 6 + i], min_diag), max_diag) / radius;
      g_state.d2e[(size_t)e * 6 + i] = d2;
      H[i * 6 + i] += d2;
    }
    if (!chol6(H, L)) { bad |= 1; continue; }
    memcpy(&g_state.L[(size_t)e * 36], L, sizeof(L));
    // Y = L^-1 [W_1 .. W_m | H_es | g_e]
    const int ncol = 6 * m + ns + 1;
    std::vector<double> Y((size_t)6 * ncol);
    std::vector<int32_t> fidx((size_t)m);
    for (int q = 0; q < m; ++q) {
      const int64_t bb = P.e_list[P.e_ptr[e] + q];
      fidx[q] = f_of[bb];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) Y[(size_t)i * ncol + 6 * q + j] = P.W[(size_t)bb * 36 + i * 6 + j];
    }
    for (int i = 0; i < 6; ++i) {
      for (int s = 0; s < ns; ++s) Y[(size_t)i * ncol + 6 * m + s] = P.Hes[((size_t)e * 6 + i) * ns + s];
      Y[(size_t)i * ncol + 6 * m + ns] = P.ge[(size_t)e * 6 + i];
    }
    fwd6(L, Y.data(), ncol, ncol);
    const double* yg = nullptr;  // last column
    (void)yg;
    // S -= Y^T Y over cell pairs (f1 <= f2 by index; shared border = cell nf)
    auto cell_update = [&](int c1, int col1, int w1, int c2, int col2, int w2, int row0, int colS) {
      // 6(or ns) x 6(or ns) block: rows = columns col1.. of Y, cols = columns col2.. of Y
      std::atomic_flag& lk = locks[(size_t)c1 * (nf + 1) + c2];
      while (lk.test_and_set(std::memory_order_acquire)) {}
      for (int i = 0; i < w1; ++i)
        for (int j = 0; j < w2; ++j) {
          double v = 0.0;
          for (int k = 0; k < 6; ++k) v += Y[(size_t)k * ncol + col1 + i] * Y[(size_t)k * ncol + col2 + j];
          S[(size_t)(row0 + i) * n_red + colS + j] -= v;
        }
      lk.clear(std::memory_order_release);
    };
    for (int q1 = 0; q1 < m; ++q1) {
      for (int q2 = 0; q2 < m; ++q2) {
        if (fidx[q2] < fidx[q1]) continue;
        if (fidx[q2] == fidx[q1] && q2 != q1 && q2 < q1) {
          // duplicate kept block seen twice in this row (rig): keep both orderings on the diagonal cell
        }
        cell_update(fidx[q1], 6 * q1, 6, fidx[q2], 6 * q2, 6, 6 * fidx[q1], 6 * fidx[q2]);
      }
      // border: kept block x shared, and rhs
      cell_update(fidx[q1], 6 * q1, 6, nf, 6 * m, ns, 6 * fidx[q1], 6 * nf);
      {
        std::atomic_flag& lk = locks[(size_t)fidx[q1] * (nf + 1) + fidx[q1]];
        while (lk.test_and_set(std::memory_order_acquire)) {}
        for (int i = 0; i < 6; ++i) {
          double v = 0.0;
          for (int k = 0; k < 6; ++k) v += Y[(size_t)k * ncol + 6 * q1 + i] * Y[(size_t)k * ncol + 6 * m + ns];
          b[6 * fidx[q1] + i] -= v;
        }
        lk.clear(std::memory_order_release);
      }
    }
    {
      std::atomic_flag& lk = locks[(size_t)nf * (nf + 1) + nf];
      while (lk.test_and_set(std::memory_order_acquire)) {}
      for (int s = 0; s < ns; ++s) {
        for (int t = s; t < ns; ++t) {
          double v = 0.0;
          for (int k = 0; k < 6; ++k) v += Y[(size_t)k * ncol + 6 * m + s] * Y[(size_t)k * ncol + 6 * m + t];
          S[(size_t)(6 * nf + s) * n_red + 6 * nf + t] -= v;
        }
        double v = 0.0;
        for (int k = 0; k < 6; ++k) v += Y[(size_t)k * ncol + 6 * m + s] * Y[(size_t)k * ncol + 6 * m + ns];
        b[6 * nf + s] -= v;
      }
      lk.clear(std::memory_order_release);
    }
  }
  // mirror the upper triangle (the diagonal cells were filled in both orders)
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n_red; ++i)
    for (int j = i + 1; j < n_red; ++j) {
      const int bi = i / 6, bj = j / 6;
      if (bi == bj && i < 6 * nf) continue;  // inside a diagonal kept cell: already full
      S[(size_t)j * n_red + i] = S[(size_t)i * n_red + j];
    }
  return bad;
}

// d_e = -(H_ee + D)^-1 (g_e + W_e d_F)
void cpu_ba_back_substitute(void* h, const uint8_t* e_const, const double* dF, double* dE) {
  Problem& P = *(Problem*)h;
  const int ns = P.n_shared, nf = P.n_f;
  const std::vector<int32_t>& f_of = P.elim_view ? P.mi : P.vi;
#pragma omp parallel for schedule(dynamic, 16)
  for (int e = 0; e < P.n_e; ++e) {
    double v[6];
    for (int i = 0; i < 6; ++i) v[i] = P.ge[(size_t)e * 6 + i];
    if (e_const && e_const[e]) { for (int i = 0; i < 6; ++i) dE[(size_t)e * 6 + i] = 0.0; continue; }
    for (int q = P.e_ptr[e]; q < P.e_ptr[e + 1]; ++q) {
      const int64_t bb = P.e_list[q];
      const double* w = &P.W[(size_t)bb * 36];
      const double* d = dF + 6 * f_of[bb];
      for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) v[i] += w[i * 6 + j] * d[j];
    }
    for (int i = 0; i < 6; ++i)
      for (int s = 0; s < ns; ++s) v[i] += P.Hes[((size_t)e * 6 + i) * ns + s] * dF[6 * nf + s];
    const double* L = &g_state.L[(size_t)e * 36];
    // solve L L^T x = v
    fwd6(L, v, 1, 1);
    for (int i = 5; i >= 0; --i) {
      double s = v[i];
      for (int k = i + 1; k < 6; ++k) s -= L[k * 6 + i] * v[k];
      v[i] = s / L[i * 6 + i];
    }
    for (int i = 0; i < 6; ++i) dE[(size_t)e * 6 + i] = -v[i];
  }
}

}  // extern "C"
