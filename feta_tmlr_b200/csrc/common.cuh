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

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// The training step is a dependent chain of a few hundred 5-20 us kernels; between two of them the GPU idles for the
// launch latency of the second.  A kernel launched with launch_chain() may start while its predecessor in the stream
// drains: its CTAs are scheduled, run their prologue (index math, shared-memory carve-up) and block in pdl_wait()
// until the predecessor has completed and flushed its writes.  Rules every such kernel follows: pdl_trigger() first,
// then no global read of producer data and NO global write before pdl_wait().  Launched without the attribute
// both calls are no-ops.
// MEASURED, and therefore opt-in (FETA_PDL=1): inside the step's CUDA graph the early-scheduled CTAs spin in
// pdl_wait() on SM slots that the graph's parallel branches (weight-gradient reductions, gradient all-reduce slices)
// would otherwise use -- ZINC step 1.49 -> 1.83 ms, PATTERN 1.67 -> 1.79 ms with it on.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

bool pdl_enabled();   // common.cu: FETA_PDL == "1"

template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// int32 exclusive scan of n elements (out may alias in); out[n] receives the total when
// write_total != 0.  `scratch` needs scan_scratch_ints(n) int32 words.
size_t scan_scratch_ints(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int write_total, int32_t* scratch,
                       cudaStream_t stream);

}  // namespace feta
