// Shared host-side helpers: error handling, device buffers, launch geometry.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <atomic>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rcc_ba.h"

namespace rcc {

struct Error : public std::runtime_error {
  int status;
  Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

#define RCC_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      throw ::rcc::Error(RCC_CUDA_ERROR, std::string(#expr) + ": " + cudaGetErrorString(_e) + \
                                             " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

#define RCC_REQUIRE(cond, status, msg)                 \
  do {                                                 \
    if (!(cond)) throw ::rcc::Error((status), (msg));  \
  } while (0)

// device buffer with explicit lifetime (owned by Problem)
template <typename T>
struct DBuf {
  T* p = nullptr;
  size_t n = 0;
  DBuf() = default;
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  ~DBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    if (count == n && p) return;
    release();
    if (count == 0) return;
    RCC_CUDA(cudaMalloc(&p, count * sizeof(T)));
    n = count;
  }
  void ensure(size_t count) {
    if (count > n) alloc(count);
  }
  void upload(const T* h, size_t count, cudaStream_t s) {
    ensure(count);
    if (count) RCC_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void upload(const std::vector<T>& h, cudaStream_t s) { upload(h.data(), h.size(), s); }
  void download(T* h, size_t count, cudaStream_t s) const {
    if (count) RCC_CUDA(cudaMemcpyAsync(h, p, count * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
  void zero(cudaStream_t s) {
    if (n) RCC_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
  }
};

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute is per device, and a process may hold
// handles on several devices (include/rcc_ba.h: "distinct handles are independent"), so the opt-in is
// remembered per (call site, current device); thread-safe.
struct SmemOptIn {
  std::atomic<uint64_t> done[4] = {};   // bit d of word d/64: set on device d
  template <typename K>
  void ensure(K kernel, size_t smem) {
    int dev = 0;
    RCC_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    std::atomic<uint64_t>& w = done[(dev >> 6) & 3];
    if (w.load(std::memory_order_acquire) & bit) return;
    RCC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    w.fetch_or(bit, std::memory_order_release);
  }
};

inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

constexpr int NUM_SMS_B200 = 148;

}  // namespace rcc
