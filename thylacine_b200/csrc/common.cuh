// Shared helpers: error plumbing, device buffers, warp reductions, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <atomic>
#include "../../include/thylacine_b200.h"

namespace thy {

// ---- error plumbing (no C++ exception crosses the ABI) ---------------------------------------------
void set_last_error(const std::string& msg);
thy_status fail(thy_status code, const std::string& msg);
// Handle-local error text: every API entry point that takes a model (or a sampler built on one) opens an ErrScope; while
// it is open, fail() also records the message in that model, where thy_model_last_error finds it from any thread.
struct ErrScope {
  thy_model_s* prev;
  explicit ErrScope(thy_model_s* m);
  ~ErrScope();
};

#define THY_CUDA_CHECK(expr)                                                                    \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      return ::thy::fail(THY_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));     \
    }                                                                                           \
  } while (0)

#define THY_TRY(expr)                   \
  do {                                  \
    thy_status _s = (expr);             \
    if (_s != THY_OK) return _s;        \
  } while (0)

extern std::atomic<uint64_t> g_kernel_launches;
inline void count_launch(uint64_t n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

thy_status require_device();  // THY_ERR_NO_DEVICE when no CUDA device is usable
int sm_count();

// ---- device buffer ----------------------------------------------------------------------------------
template <typename T>
struct DevBuf {
  T* ptr = nullptr;
  size_t count = 0;
  uint64_t* realloc_counter = nullptr;   // bumped whenever the buffer moves (captured CUDA graphs hold the old pointer)
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    count = 0;
  }
  // grow-only
  thy_status reserve(size_t n) {
    if (n <= count) return THY_OK;
    release();
    if (realloc_counter) ++*realloc_counter;
    THY_CUDA_CHECK(cudaMalloc(&ptr, n * sizeof(T)));
    count = n;
    return THY_OK;
  }
  thy_status upload(const std::vector<T>& h, cudaStream_t s = 0) {
    THY_TRY(reserve(h.size() ? h.size() : 1));
    if (!h.empty()) THY_CUDA_CHECK(cudaMemcpyAsync(ptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return THY_OK;
  }
};

// ---- warp helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double quad_sum(double v) {  // over the 4 lanes sharing lane/4
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline long long round_up(long long a, long long b) { return (a + b - 1) / b * b; }

}  // namespace thy
