// TEST-ONLY driver: builds a ceres::Problem (tests/ceres_mock) through include/rcc_ceres_adapter.h from a scene
// file written by tests/test_ceres_adapter.py, evaluates it the way Ceres would and writes residuals + Jacobians
// back for comparison with the oracle.  Links librcc_ba.so; needs a GPU to run, not to compile.
//   scene file (doubles unless noted): header int64 {n_views, n_markers, n_obs}; intr[4] dist[5]; views[6 nv];
//   markers[6 nm]; sizes[nm]; int32 view_idx[n]; int32 marker_idx[n]; pixels[8 n]
#include <cstdio>
#include <cstdlib>

#include "rcc_ceres_adapter.h"

template <typename T>
static std::vector<T> rd(FILE* f, size_t n) {
  std::vector<T> v(n);
  if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
  return v;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: adapter_driver scene.bin out.bin\n"); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  auto hdr = rd<int64_t>(f, 3);
  const int64_t nv = hdr[0], nm = hdr[1], n = hdr[2];
  auto intr = rd<double>(f, 4), dist = rd<double>(f, 5), views = rd<double>(f, 6 * nv), markers = rd<double>(f, 6 * nm),
       sizes = rd<double>(f, nm);
  auto vi = rd<int32_t>(f, n), mi = rd<int32_t>(f, n);
  auto pix = rd<double>(f, 8 * n);
  fclose(f);

  ceres::Problem problem;
  rcc_ceres::GpuBatch batch(intr.data(), dist.data());
  for (int64_t i = 0; i < n; ++i) {
    double* v = &views[6 * vi[i]];
    double* m = &markers[6 * mi[i]];
    problem.AddResidualBlock(batch.AddTag(v, m, sizes[mi[i]], &pix[8 * i]), nullptr, intr.data(), dist.data(), v, m);
  }
  problem.SetParameterBlockConstant(&markers[0]);        // the world tag, camera_pose.cpp:71-80
  batch.Finalize(0);

  std::vector<double> r0, r1, J1, r2, J2;
  bool ok = problem.Evaluate(&batch, false, true, &r0, nullptr);              // residuals only
  ok = ok && problem.Evaluate(&batch, true, false, &r1, &J1);                 // same point, now with Jacobians
  // move the point the way a solver step would (in the user's arrays), evaluate again
  for (auto& x : views) x += 1e-3;
  intr[0] *= 1.001;
  ok = ok && problem.Evaluate(&batch, true, true, &r2, &J2);
  FILE* o = fopen(argv[2], "wb");
  int64_t meta[4] = {ok ? 1 : 0, (int64_t)r1.size(), (int64_t)J1.size(), batch.evaluations()};
  fwrite(meta, sizeof(int64_t), 4, o);
  fwrite(r0.data(), 8, r0.size(), o);
  fwrite(r1.data(), 8, r1.size(), o);
  fwrite(J1.data(), 8, J1.size(), o);
  fwrite(r2.data(), 8, r2.size(), o);
  fwrite(J2.data(), 8, J2.size(), o);
  fclose(o);
  return ok ? 0 : 1;
}
