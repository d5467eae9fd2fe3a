// Library-wide state (error text, launch counter) and a small int32 device scan.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace feta {

std::atomic<int64_t> g_launch_count{0};

char* last_error_buf() {
  static thread_local char buf[512] = "ok";
  return buf;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("FETA_PDL");     // opt-in: measured SLOWER inside the step graph (see common.cuh)
    return e != nullptr && e[0] == '1';
  }();
  return on;
}

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
}

// ---- exclusive scan: 256 threads x 4 items per block, recursive over block sums ----------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 4;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_blocks_kernel(const int32_t* __restrict__ in,
                                                                   int32_t* out, int64_t n,
                                                                   int32_t* __restrict__ block_sums) {
  __shared__ int32_t warp_tot[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  int32_t tsum = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    tsum += v[i];
  }
  // inclusive warp scan of thread sums
  int32_t inc = tsum;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  int32_t warp_off = 0;
  for (int w = 0; w < warp; ++w) warp_off += warp_tot[w];
  int32_t run = warp_off + inc - tsum;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
  if (threadIdx.x == kScanThreads - 1) block_sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(kScanThreads) scan_add_offsets_kernel(int32_t* out, int64_t n,
                                                                        const int32_t* __restrict__ offs) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  const int32_t o = offs[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) out[base + i] += o;
}

__global__ void copy_word_kernel(const int32_t* src, int32_t* dst) { *dst = *src; }

size_t scan_scratch_ints(int64_t n) {
  size_t tot = 0;
  while (true) {
    int64_t nb = ceil_div(n > 0 ? n : 1, kScanTile);
    tot += (size_t)nb + 1;
    if (nb <= 1) break;
    n = nb;
  }
  return tot + 8;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int write_total, int32_t* scratch,
                       cudaStream_t stream) {
  if (n <= 0) {
    if (write_total) FETA_CUDA(cudaMemsetAsync(out, 0, sizeof(int32_t), stream));
    return FETA_OK;
  }
  const int64_t nb = ceil_div(n, kScanTile);
  scan_blocks_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(in, out, n, scratch);
  FETA_LAUNCH_CHECK();
  if (nb > 1) {
    int rc = exclusive_scan_i32(scratch, scratch, nb, 1, scratch + nb + 1, stream);
    if (rc != FETA_OK) return rc;
    scan_add_offsets_kernel<<<(unsigned)nb, kScanThreads, 0, stream>>>(out, n, scratch);
    FETA_LAUNCH_CHECK();
  }
  if (write_total) {
    copy_word_kernel<<<1, 1, 0, stream>>>(scratch + (nb > 1 ? nb : 0), out + n);
    FETA_LAUNCH_CHECK();
  }
  return FETA_OK;
}

}  // namespace feta

extern "C" {
int feta_version(void) { return 100; }
const char* feta_last_error_string(void) { return feta::last_error_buf(); }
int64_t feta_launch_count(void) { return feta::g_launch_count.load(); }
}
