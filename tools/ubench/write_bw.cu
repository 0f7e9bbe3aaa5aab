// Microbenchmark: pure HBM write bandwidth on B200 (the ceiling of the write-bound K1 kernel).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_bw write_bw.cu && ./write_bw
#include <cstdio>
#include <cuda_runtime.h>

__global__ void st128(double2* p, size_t n) {
  const double2 v = make_double2(1.0, 2.0);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void st128_cs(double2* p, size_t n) {   // streaming (evict-first) stores
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    __stcs(p + i, make_double2(1.0, 2.0));
}
// TMA bulk copy smem -> global, 2 KB per copy, one elected thread per warp
__global__ void bulk(double* p, size_t n_chunks) {
  __shared__ __align__(128) double buf[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 256; i += 32) buf[warp][i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  const size_t nw = (size_t)gridDim.x * 8;
  for (size_t c = (size_t)blockIdx.x * 8 + warp; c < n_chunks; c += nw) {
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 2048;" ::"l"(p + c * 256),
                   "r"((unsigned)__cvta_generic_to_shared(buf[warp])) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__global__ void copy128(const double2* __restrict__ a, double2* __restrict__ b, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

int main() {
  const size_t bytes = (size_t)2 << 30;
  double *a, *b; cudaMalloc(&a, bytes); cudaMalloc(&b, bytes);
  cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](const char* name, auto launch, double moved) {
    float best = 1e9f;
    for (int r = 0; r < 6; ++r) {
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (r > 0 && ms < best) best = ms;
    }
    printf("%-40s %8.3f ms  %8.1f GB/s  (%s)\n", name, best, moved / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  const size_t n16 = bytes / 16;
  run("STG.128 grid-stride, 148x8 CTAs", [&] { st128<<<148 * 8, 256>>>((double2*)a, n16); }, (double)bytes);
  run("STG.128 grid-stride, 148x32 CTAs", [&] { st128<<<148 * 32, 256>>>((double2*)a, n16); }, (double)bytes);
  run("STG.128 .cs streaming", [&] { st128_cs<<<148 * 16, 256>>>((double2*)a, n16); }, (double)bytes);
  run("cp.async.bulk smem->global 2 KB", [&] { bulk<<<148 * 8, 256>>>(a, bytes / 2048); }, (double)bytes);
  run("cudaMemsetAsync", [&] { cudaMemsetAsync(a, 0, bytes); }, (double)bytes);
  run("copy LDG.128/STG.128 (read+write bytes)", [&] { copy128<<<148 * 16, 256>>>((const double2*)a, (double2*)b, n16); }, 2.0 * bytes);
  return 0;
}
