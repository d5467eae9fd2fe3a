// Shared helpers for libfeta_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/feta_b200.h"

namespace feta {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// thread-local last-error text (feta_last_error_string)
char* last_error_buf();
void set_last_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launch_count;

inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define FETA_REQUIRE(cond, ...)                      \
  do {                                               \
    if (!(cond)) {                                   \
      ::feta::set_last_error(__VA_ARGS__);           \
      return FETA_EINVAL;                            \
    }                                                \
  } while (0)

#define FETA_CUDA(call)                                                                      \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      ::feta::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                  \
                             cudaGetErrorString(e_));                                        \
      return FETA_ECUDA;                                                                     \
    }                                                                                        \
  } while (0)

// after a <<<>>> launch: catches bad configs without synchronising
#define FETA_LAUNCH_CHECK()                                                                  \
  do {                                                                                       \
    ::feta::count_launch();                                                                  \
    cudaError_t e_ = cudaGetLastError();                                                     \
    if (e_ != cudaSuccess) {                                                                 \
      ::feta::set_last_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,              \
                             cudaGetErrorString(e_));                                        \
      return FETA_ECUDA;                                                                     \
    }                                                                                        \
  } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// bump allocator over a caller-provided workspace
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t n) : base((char*)p), cap(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T), 256);
    if (off + bytes > cap) return nullptr;
    T* r = (T*)(base + off);
    off += bytes;
    return r;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// int32 exclusive scan of n elements (out may alias in); out[n] receives the total when
// write_total != 0.  `scratch` needs scan_scratch_ints(n) int32 words.
size_t scan_scratch_ints(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int write_total, int32_t* scratch,
                       cudaStream_t stream);

}  // namespace feta
