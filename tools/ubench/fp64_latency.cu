// Microbenchmark: dependent-issue latency and per-warp throughput of FP64 ops on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chain(double* out, int iters, long long* cycles) {
  double a[CHAINS];
  for (int c = 0; c < CHAINS; ++c) a[c] = threadIdx.x * 1e-9 + c;
  const double m = 1.0000001, k = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u)
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) a[c] = fma(a[c], m, k);
  }
  long long t1 = clock64();
  double s = 0;
  for (int c = 0; c < CHAINS; ++c) s += a[c];
  if (s == 123.456) out[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
__global__ void ddiv_chain(double* out, int iters, long long* cycles) {
  double a = 1.0 + threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) a = 1.0 / (a + 0.5);
  }
  long long t1 = clock64();
  if (a == 123.456) out[0] = a;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}
__global__ void lds_chain(double* out, int iters, long long* cycles) {
  __shared__ int idx[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) idx[i] = (i + 1) & 255;
  __syncthreads();
  int p = threadIdx.x & 255;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) p = idx[p];
  }
  long long t1 = clock64();
  if (p == 12345) out[0] = p;
  if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  auto report = [&](const char* name, double ops) {
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-46s %8.2f cycles/op\n", name, (double)h / ops);
  };
  dfma_chain<1><<<1, 32>>>(out, iters, cyc); report("DFMA dependent chain, 1 warp", iters * 16.0);
  dfma_chain<2><<<1, 32>>>(out, iters, cyc); report("DFMA 2 chains, 1 warp (per DFMA)", iters * 32.0);
  dfma_chain<4><<<1, 32>>>(out, iters, cyc); report("DFMA 4 chains, 1 warp (per DFMA)", iters * 64.0);
  dfma_chain<8><<<1, 32>>>(out, iters, cyc); report("DFMA 8 chains, 1 warp (per DFMA)", iters * 128.0);
  dfma_chain<8><<<1, 128>>>(out, iters, cyc); report("DFMA 8 chains, 4 warps = 1/SMSP (per DFMA/warp)", iters * 128.0);
  dfma_chain<8><<<1, 256>>>(out, iters, cyc); report("DFMA 8 chains, 8 warps = 2/SMSP (per DFMA/warp)", iters * 128.0);
  dfma_chain<1><<<1, 256>>>(out, iters, cyc); report("DFMA 1 chain, 8 warps = 2/SMSP (per DFMA/warp)", iters * 16.0);
  dfma_chain<1><<<1, 512>>>(out, iters, cyc); report("DFMA 1 chain, 16 warps = 4/SMSP (per DFMA/warp)", iters * 16.0);
  ddiv_chain<<<1, 32>>>(out, iters, cyc); report("1/(x+0.5) dependent chain, 1 warp", iters * 16.0);
  lds_chain<<<1, 32>>>(out, iters, cyc); report("LDS.32 pointer chase, 1 warp", iters * 16.0);
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
