// Batched PnP initialiser -- the job cv::solvePnP(obj_pts, img_pts, K, dist, rvec,
// tvec, false, CV_ITERATIVE) does once per tag per frame at
// /root/reference/real_preprocessing/src/camera_pose.cpp:163 (tagTcam, :132-173):
// minimise the reprojection error of the four tag corners over cam_T_tag (6 dof).
//
// One thread per tag.  Like OpenCV's iterative solver for planar targets it
// starts from the homography between the tag plane and the undistorted
// normalised image points and refines with Levenberg-Marquardt; the cost and
// its analytic Jacobian are the same model.cuh functions the bundle adjustment
// uses (view = identity, marker = cam_T_tag).
#include "common.cuh"
#include "kernels.h"
#include "model.cuh"

namespace rcc {

// solve the n x n system A x = b in place (partial pivoting); returns false if singular
template <int N>
__device__ bool solve_dense(double (&A)[N][N], double (&b)[N]) {
  for (int c = 0; c < N; ++c) {
    int piv = c;
    double best = fabs(A[c][c]);
    for (int r = c + 1; r < N; ++r)
      if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); piv = r; }
    if (!(best > 1e-300)) return false;
    if (piv != c) {
      for (int k = 0; k < N; ++k) { const double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
      const double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    const double inv = 1.0 / A[c][c];
    for (int r = c + 1; r < N; ++r) {
      const double f = A[r][c] * inv;
      if (f != 0.0) {
        for (int k = c; k < N; ++k) A[r][k] -= f * A[c][k];
        b[r] -= f * b[c];
      }
    }
  }
  for (int r = N - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < N; ++k) s -= A[r][k] * b[k];
    b[r] = s / A[r][r];
  }
  return true;
}

// rotation matrix (row-major) -> Rodrigues vector
__device__ void rot_to_rvec(const double* R, double* r) {
  const double wx = 0.5 * (R[7] - R[5]), wy = 0.5 * (R[2] - R[6]), wz = 0.5 * (R[3] - R[1]);
  const double s = sqrt(wx * wx + wy * wy + wz * wz);
  const double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
  const double t = atan2(s, c);
  if (s > 1e-9) {
    const double k = t / s;
    r[0] = wx * k; r[1] = wy * k; r[2] = wz * k;
  } else if (c > 0.0) {
    r[0] = wx; r[1] = wy; r[2] = wz;
  } else {
    // angle ~ pi: axis from the diagonal of (R + I)/2
    const double xx = 0.5 * (R[0] + 1.0), yy = 0.5 * (R[4] + 1.0), zz = 0.5 * (R[8] + 1.0);
    double ax = sqrt(fmax(xx, 0.0)), ay = sqrt(fmax(yy, 0.0)), az = sqrt(fmax(zz, 0.0));
    if (ax >= ay && ax >= az) { ay = copysign(ay, R[1] + R[3]); az = copysign(az, R[2] + R[6]); }
    else if (ay >= az) { ax = copysign(ax, R[1] + R[3]); az = copysign(az, R[5] + R[7]); }
    else { ax = copysign(ax, R[2] + R[6]); ay = copysign(ay, R[5] + R[7]); }
    r[0] = ax * t; r[1] = ay * t; r[2] = az * t;
  }
}

// cost and (optionally) J^T J, J^T r of the 8 residuals wrt pose p (rvec,t)
template <bool WANT_J>
__device__ double pnp_cost(const double* p, const double* sh, double hs, const double* pix, double* H, double* g,
                           bool* ok) {
  double idx[POSEX], px[POSEX];
  const double zero[6] = {0, 0, 0, 0, 0, 0};
  expand_pose(zero, idx);
  expand_marker_pose(p, px);   // the tag pose plays the marker role: its Jr slot holds R Jr
  BlockGeom<false> geo;
  block_geometry<false>(idx, px, nullptr, geo);
  double cost = 0.0;
  if (WANT_J) {
    for (int i = 0; i < 36; ++i) H[i] = 0.0;
    for (int i = 0; i < 6; ++i) g[i] = 0.0;
  }
  *ok = true;
  for (int k = 0; k < 4; ++k) {
    double ox, oy;
    corner_xy(k, hs, ox, oy);
    CornerRows<false> c;
    eval_corner<false, WANT_J>(geo, sh, ox, oy, pix[2 * k], pix[2 * k + 1], c);
    if (!(c.depth > 0.0) || !isfinite(c.r[0]) || !isfinite(c.r[1])) *ok = false;
    cost += c.r[0] * c.r[0] + c.r[1] * c.r[1];
    if (WANT_J) {
      for (int i = 0; i < 2; ++i)
        for (int a = 0; a < 6; ++a) {
          g[a] += c.jm[i][a] * c.r[i];
          for (int b = a; b < 6; ++b) H[a * 6 + b] += c.jm[i][a] * c.jm[i][b];
        }
    }
  }
  return 0.5 * cost;
}

__global__ void __launch_bounds__(128) pnp_kernel(int64_t n, const double* __restrict__ shared9,
                                                  const double* __restrict__ sizes, const double* __restrict__ pixels,
                                                  double* __restrict__ poses, double* __restrict__ final_cost,
                                                  int max_iter) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double sh[9], pix[8], p[6];
  for (int k = 0; k < 9; ++k) sh[k] = shared9[k];
  for (int k = 0; k < 8; ++k) pix[k] = pixels[i * 8 + k];
  for (int k = 0; k < 6; ++k) p[k] = poses[i * 6 + k];
  const double hs = 0.5 * sizes[i];
  bool have_guess = false;
  for (int k = 0; k < 6; ++k) have_guess |= (p[k] != 0.0);

  if (!have_guess) {
    // ---- undistorted normalised image points (fixed-point iteration, like cv::undistortPoints)
    double xn[4], yn[4];
    for (int k = 0; k < 4; ++k) {
      const double x0 = (pix[2 * k] - sh[2]) / sh[0], y0 = (pix[2 * k + 1] - sh[3]) / sh[1];
      double x = x0, y = y0;
      for (int it = 0; it < 20; ++it) {
        const double r2 = x * x + y * y;
        const double rad = 1.0 + r2 * (sh[4] + r2 * (sh[5] + r2 * sh[8]));
        const double dx = 2.0 * sh[6] * x * y + sh[7] * (r2 + 2.0 * x * x);
        const double dy = sh[6] * (r2 + 2.0 * y * y) + 2.0 * sh[7] * x * y;
        x = (x0 - dx) / rad;
        y = (y0 - dy) / rad;
      }
      xn[k] = x; yn[k] = y;
    }
    // ---- homography (X,Y,1) -> (x,y,1) from the 4 corners, h33 = 1
    double A[8][8], b[8];
    for (int k = 0; k < 4; ++k) {
      double X, Y;
      corner_xy(k, hs, X, Y);
      const double r0[8] = {X, Y, 1, 0, 0, 0, -xn[k] * X, -xn[k] * Y};
      const double r1[8] = {0, 0, 0, X, Y, 1, -yn[k] * X, -yn[k] * Y};
      for (int c = 0; c < 8; ++c) { A[2 * k][c] = r0[c]; A[2 * k + 1][c] = r1[c]; }
      b[2 * k] = xn[k]; b[2 * k + 1] = yn[k];
    }
    bool okh = solve_dense<8>(A, b);
    double h1[3] = {b[0], b[3], b[6]}, h2[3] = {b[1], b[4], b[7]}, h3[3] = {b[2], b[5], 1.0};
    if (okh) {
      // H = [r1 r2 t] / t_z up to scale (h33 = 1 and t_z > 0 for a tag in front of the camera)
      const double n1 = sqrt(h1[0] * h1[0] + h1[1] * h1[1] + h1[2] * h1[2]);
      const double n2 = sqrt(h2[0] * h2[0] + h2[1] * h2[1] + h2[2] * h2[2]);
      const double lam = 2.0 / (n1 + n2);
      double r1v[3], r2v[3], r3v[3];
      for (int k = 0; k < 3; ++k) r1v[k] = h1[k] / n1;
      const double d12 = r1v[0] * h2[0] + r1v[1] * h2[1] + r1v[2] * h2[2];
      for (int k = 0; k < 3; ++k) r2v[k] = h2[k] - d12 * r1v[k];   // Gram-Schmidt
      const double nn = sqrt(r2v[0] * r2v[0] + r2v[1] * r2v[1] + r2v[2] * r2v[2]);
      for (int k = 0; k < 3; ++k) r2v[k] /= nn;
      r3v[0] = r1v[1] * r2v[2] - r1v[2] * r2v[1];
      r3v[1] = r1v[2] * r2v[0] - r1v[0] * r2v[2];
      r3v[2] = r1v[0] * r2v[1] - r1v[1] * r2v[0];
      const double R[9] = {r1v[0], r2v[0], r3v[0], r1v[1], r2v[1], r3v[1], r1v[2], r2v[2], r3v[2]};
      rot_to_rvec(R, p);
      p[3] = h3[0] * lam; p[4] = h3[1] * lam; p[5] = h3[2] * lam;
    } else {
      p[0] = p[1] = p[2] = 0.0; p[3] = p[4] = 0.0; p[5] = 1.0;
    }
  }

  // ---- Levenberg-Marquardt on the 8 residuals
  double H[36], g[6];
  bool ok;
  double cost = pnp_cost<true>(p, sh, hs, pix, H, g, &ok);
  double lambda = 1e-3;
  for (int it = 0; it < max_iter; ++it) {
    double A6[6][6], rhs[6];
    for (int a = 0; a < 6; ++a) {
      for (int c = 0; c < 6; ++c) A6[a][c] = (c >= a) ? H[a * 6 + c] : H[c * 6 + a];
      A6[a][a] += lambda * fmax(H[a * 6 + a], 1e-12);
      rhs[a] = -g[a];
    }
    if (!solve_dense<6>(A6, rhs)) { lambda *= 10.0; continue; }
    double q[6];
    double dn = 0.0, pn = 0.0;
    for (int a = 0; a < 6; ++a) { q[a] = p[a] + rhs[a]; dn += rhs[a] * rhs[a]; pn += p[a] * p[a]; }
    bool ok2;
    const double c2 = pnp_cost<false>(q, sh, hs, pix, nullptr, nullptr, &ok2);
    // Accept on "not worse up to rounding": along the weakly determined depth direction of a small tag the cost is
    // flat to 1e-16 relative while the step is still 1e-8, so convergence is declared on the step alone.
    // Below a relative step of 1e-5 the Gauss-Newton model is exact to rounding while the cost comparison is pure
    // noise (it would stall the iteration ~1e-8 short of the minimiser): such steps are taken on trust.
    const bool tiny = dn <= 1e-10 * (pn + 1e-10);
    if (ok2 && (tiny || c2 <= cost * (1.0 + 4e-16))) {
      for (int a = 0; a < 6; ++a) p[a] = q[a];
      cost = pnp_cost<true>(p, sh, hs, pix, H, g, &ok);
      lambda = fmax(lambda * 0.1, 1e-12);
      if (dn <= 1e-26 * (pn + 1e-26)) break;
    } else {
      lambda *= 10.0;
      if (lambda > 1e12) break;
    }
  }
  for (int k = 0; k < 6; ++k) poses[i * 6 + k] = p[k];
  if (final_cost) final_cost[i] = ok ? cost : -1.0;
}

}  // namespace rcc

extern "C" int rcc_pnp_batch(int32_t device, int64_t n, const double* shared9, const double* sizes,
                             const double* pixels, double* cam_T_tag, double* final_cost, int32_t max_iterations) {
  using namespace rcc;
  if (n < 0 || !shared9 || (n > 0 && (!sizes || !pixels || !cam_T_tag))) return RCC_BAD_ARG;
  if (n == 0) return RCC_OK;
  try {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      throw Error(RCC_CUDA_ERROR, "no CUDA device (this library has no CPU fallback)");
    RCC_CUDA(cudaSetDevice(device));
    DBuf<double> dsh, dsz, dpx, dpose, dcost;
    cudaStream_t s;
    RCC_CUDA(cudaStreamCreate(&s));
    dsh.upload(shared9, 9, s);
    dsz.upload(sizes, (size_t)n, s);
    dpx.upload(pixels, (size_t)n * 8, s);
    dpose.upload(cam_T_tag, (size_t)n * 6, s);
    dcost.alloc((size_t)n);
    pnp_kernel<<<ceil_div(n, 128), 128, 0, s>>>(n, dsh.p, dsz.p, dpx.p, dpose.p, dcost.p,
                                                  max_iterations > 0 ? max_iterations : 30);
    RCC_CUDA(cudaGetLastError());
    dpose.download(cam_T_tag, (size_t)n * 6, s);
    if (final_cost) dcost.download(final_cost, (size_t)n, s);
    RCC_CUDA(cudaStreamSynchronize(s));
    cudaStreamDestroy(s);
  } catch (const Error& e) {
    fprintf(stderr, "[rcc_pnp_batch] %s\n", e.what());
    return e.status;
  }
  return RCC_OK;
}
