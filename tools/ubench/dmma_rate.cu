// Microbenchmark: throughput of FP64 tensor-core MMA (mma.sync m8n8k4 / m16n8k8 f64) vs scalar DFMA on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_rate dmma_rate.cu && ./dmma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int CHAINS>
__global__ void __launch_bounds__(256) k884(double* out, int iters) {
  double c[CHAINS][2];
  for (int i = 0; i < CHAINS; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; }
  const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) mma884(c[i][0], c[i][1], a, b);
  double s = 0;
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
template <int CHAINS>
__global__ void __launch_bounds__(256) k1688(double* out, int iters) {
  double c[CHAINS][4];
  for (int i = 0; i < CHAINS; ++i) for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 1e-9 + i + j;
  const double a[4] = {1.0 + threadIdx.x * 1e-12, 1e-3, 2e-3, 3e-3}, b[2] = {1e-9, 2e-9};
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) mma1688(c[i], a, b);
  double s = 0;
  for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}
template <int CHAINS>
__global__ void __launch_bounds__(256) kdfma(double* out, int iters) {
  double c[CHAINS];
  for (int i = 0; i < CHAINS; ++i) c[i] = threadIdx.x * 1e-9 + i;
  const double a = 1.0000001, b = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
  double s = 0;
  for (int i = 0; i < CHAINS; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}

int main() {
  double* out; cudaMalloc(&out, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  const int grid = p.multiProcessorCount * 8, iters = 4000;
  auto run = [&](const char* name, auto launch, double fma_per_thread_iter) {
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)grid * 256 * iters * fma_per_thread_iter;
    printf("%-44s %8.3f ms  %7.2f TFLOP/s\n", name, ms, 2.0 * fmas / (ms * 1e-3) / 1e12);
  };
  // per thread per mma: m8n8k4 = 256 FMA / 32 lanes = 8 ; m16n8k8 = 1024 / 32 = 32
  run("DFMA scalar, 8 chains", [&] { kdfma<8><<<grid, 256>>>(out, iters); }, 4.0 * 8);
  run("DMMA m8n8k4, 1 chain/warp", [&] { k884<1><<<grid, 256>>>(out, iters); }, 4.0 * 1 * 8);
  run("DMMA m8n8k4, 4 chains/warp", [&] { k884<4><<<grid, 256>>>(out, iters); }, 4.0 * 4 * 8);
  run("DMMA m8n8k4, 8 chains/warp", [&] { k884<8><<<grid, 256>>>(out, iters); }, 4.0 * 8 * 8);
  run("DMMA m16n8k8, 1 chain/warp", [&] { k1688<1><<<grid, 256>>>(out, iters); }, 4.0 * 1 * 32);
  run("DMMA m16n8k8, 4 chains/warp", [&] { k1688<4><<<grid, 256>>>(out, iters); }, 4.0 * 4 * 32);
  // single-warp latency
  for (int w = 0; w < 2; ++w) {
    cudaEventRecord(e0);
    if (w == 0) k884<1><<<1, 32>>>(out, iters); else k1688<1><<<1, 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%s dependent latency: %.1f ns per mma\n", w == 0 ? "m8n8k4" : "m16n8k8", ms * 1e6 / (iters * 4.0));
  }
  return 0;
}
