// A5: scatter / pool / pack ops (see include/feta_b200.h).
//
// Replaces the torch_scatter / PyG / advanced-indexing paths of transformer/models.py:
// :177-185 + :347 (head stacking + packed gather), :200-202 (un-stack + scatter back into a
// zero-filled padded tensor), :283 (global_mean_pool), :586-595 (GlobalAvg1D), :1070-1071
// (cls_output[~masks]).  All are pure data movement / segmented sums: one element per thread,
// consecutive threads on consecutive channels (coalesced), no atomics (segments are sorted).
#include "common.cuh"

namespace feta {

constexpr int kSegThreads = 256;

static inline unsigned seg_grid(int64_t n) {
  int64_t b = ceil_div(n > 0 ? n : 1, kSegThreads);
  const int64_t cap = (int64_t)kNumSMs * 32;
  return (unsigned)(b < cap ? b : cap);
}

// x[(h*N + i), c] <-> o_heads[fi[i,0], fi[i,1], h, c]
template <bool kBackward>
__global__ void __launch_bounds__(kSegThreads) pack_heads_kernel(const float* __restrict__ src,
                                                                const int64_t* __restrict__ fi,
                                                                float* __restrict__ dst, int64_t N, int nmax, int H,
                                                                int dh) {
  const int64_t total = (int64_t)H * N * dh;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % dh);
    const int64_t row = idx / dh;
    const int h = (int)(row / N);
    const int64_t i = row - (int64_t)h * N;
    const int64_t b = fi[2 * i], n = fi[2 * i + 1];
    const int64_t padded = ((b * nmax + n) * H + h) * dh + c;
    if (!kBackward)
      dst[idx] = src[padded];
    else
      dst[padded] = src[idx];
  }
}

// y[(h*N + i), c] <-> out[fi[i,1], fi[i,0], h*dh + c]   (out is [Nmax, B, H*dh])
template <bool kBackward>
__global__ void __launch_bounds__(kSegThreads) unpack_heads_kernel(const float* __restrict__ src,
                                                                  const int64_t* __restrict__ fi,
                                                                  float* __restrict__ dst, int64_t N, int B, int H,
                                                                  int dh) {
  const int64_t total = (int64_t)H * N * dh;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % dh);
    const int64_t row = idx / dh;
    const int h = (int)(row / N);
    const int64_t i = row - (int64_t)h * N;
    const int64_t b = fi[2 * i], n = fi[2 * i + 1];
    const int64_t padded = ((n * B + b) * H + h) * dh + c;
    if (!kBackward)
      dst[padded] = src[idx];
    else
      dst[idx] = src[padded];
  }
}

__global__ void __launch_bounds__(kSegThreads) segment_mean_fwd_kernel(const float* __restrict__ x,
                                                                      const int32_t* __restrict__ graph_ptr,
                                                                      float* __restrict__ out, int C) {
  const int g = blockIdx.x;
  const int lo = graph_ptr[g], hi = graph_ptr[g + 1];
  const float inv = 1.0f / (float)(hi - lo > 0 ? hi - lo : 1);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.0f;
    for (int r = lo; r < hi; ++r) acc += x[(size_t)r * C + c];
    out[(size_t)g * C + c] = acc * inv;
  }
}

__global__ void __launch_bounds__(kSegThreads) segment_mean_bwd_kernel(const float* __restrict__ d_out,
                                                                      const int32_t* __restrict__ graph_ptr,
                                                                      float* __restrict__ dx, int C) {
  const int g = blockIdx.x;
  const int lo = graph_ptr[g], hi = graph_ptr[g + 1];
  const float inv = 1.0f / (float)(hi - lo > 0 ? hi - lo : 1);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float v = d_out[(size_t)g * C + c] * inv;
    for (int r = lo; r < hi; ++r) dx[(size_t)r * C + c] = v;
  }
}

// GlobalAvg1D: out[b, c] = sum_{n real} x[b, n, c] / count_b
__global__ void __launch_bounds__(kSegThreads) masked_mean_fwd_kernel(const float* __restrict__ x, int64_t sb,
                                                                     int64_t sn, const uint8_t* __restrict__ mask,
                                                                     float* __restrict__ out, int nmax, int C) {
  const int b = blockIdx.x;
  const uint8_t* mk = mask + (size_t)b * nmax;
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int loc = 0;
  for (int n = threadIdx.x; n < nmax; n += blockDim.x) loc += (mk[n] == 0);
  if (loc) atomicAdd(&s_cnt, loc);
  __syncthreads();
  const float cnt = (float)s_cnt;
  // thread (slice, c): rows slice, slice + nsl, ... of column c (coalesced across c), slices folded in shared memory
  // in a fixed order
  extern __shared__ float red[];                       // [blockDim.x]
  const int nsl = C <= (int)blockDim.x ? (int)blockDim.x / C : 1;
  for (int c0 = 0; c0 < C; c0 += blockDim.x) {
    const int c = c0 + (int)threadIdx.x % (C < (int)blockDim.x ? C : (int)blockDim.x);
    const int slice = C < (int)blockDim.x ? (int)threadIdx.x / C : 0;
    float acc = 0.0f;
    if (c < C && slice < nsl)
      for (int n = slice; n < nmax; n += nsl)
        if (mk[n] == 0) acc += x[(int64_t)b * sb + (int64_t)n * sn + c];
    red[threadIdx.x] = acc;
    __syncthreads();
    if (slice == 0 && c < C) {
      float t = acc;
      for (int k = 1; k < nsl; ++k) t += red[k * C + (c - c0)];
      out[(size_t)b * C + c] = t / cnt;  // the reference divides by mask.sum() un-clamped
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kSegThreads) masked_mean_bwd_kernel(const float* __restrict__ d_out,
                                                                     const uint8_t* __restrict__ mask,
                                                                     float* __restrict__ dx, int nmax, int C) {
  const int b = blockIdx.x;
  const uint8_t* mk = mask + (size_t)b * nmax;
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  int loc = 0;
  for (int n = threadIdx.x; n < nmax; n += blockDim.x) loc += (mk[n] == 0);
  if (loc) atomicAdd(&s_cnt, loc);
  __syncthreads();
  const float cnt = (float)s_cnt;
  for (int idx = threadIdx.x; idx < nmax * C; idx += blockDim.x) {
    const int n = idx / C, c = idx - n * C;
    dx[((size_t)b * nmax + n) * C + c] = mk[n] == 0 ? d_out[(size_t)b * C + c] / cnt : 0.0f;
  }
}

template <bool kScatter>
__global__ void __launch_bounds__(kSegThreads) gather_rows_kernel(const float* __restrict__ src,
                                                                 const int64_t* __restrict__ fi,
                                                                 float* __restrict__ dst, int64_t sb, int64_t sn,
                                                                 int64_t N, int C) {
  const int64_t total = N * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / C;
    const int c = (int)(idx - i * C);
    const int64_t padded = fi[2 * i] * sb + fi[2 * i + 1] * sn + c;
    if (!kScatter)
      dst[idx] = src[padded];
    else
      dst[padded] = src[idx];
  }
}

}  // namespace feta

using namespace feta;

extern "C" int feta_pack_heads(const float* o_heads, const int64_t* fi, float* x, int64_t N, int B, int nmax, int H,
                               int dh, void* stream_) {
  (void)B;
  if (N == 0) return FETA_OK;
  FETA_REQUIRE(o_heads && fi && x && N > 0 && H >= 1 && dh >= 1, "pack_heads: bad argument");
  pack_heads_kernel<false><<<seg_grid((int64_t)H * N * dh), kSegThreads, 0, (cudaStream_t)stream_>>>(o_heads, fi, x, N,
                                                                                                     nmax, H, dh);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_pack_heads_bwd(const float* dx, const int64_t* fi, float* d_o_heads, int64_t N, int B, int nmax,
                                   int H, int dh, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(d_o_heads || (size_t)B * nmax == 0, "pack_heads_bwd: NULL output");
  FETA_CUDA(cudaMemsetAsync(d_o_heads, 0, (size_t)B * nmax * H * dh * sizeof(float), st));
  if (N == 0) return FETA_OK;
  FETA_REQUIRE(dx && fi, "pack_heads_bwd: NULL pointer argument");
  pack_heads_kernel<true><<<seg_grid((int64_t)H * N * dh), kSegThreads, 0, st>>>(dx, fi, d_o_heads, N, nmax, H, dh);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_unpack_heads(const float* y, const int64_t* fi, float* out, int64_t N, int B, int nmax, int H,
                                 int dh, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(out || (size_t)B * nmax == 0, "unpack_heads: NULL output");
  FETA_CUDA(cudaMemsetAsync(out, 0, (size_t)B * nmax * H * dh * sizeof(float), st));  // torch.zeros, models.py:201
  if (N == 0) return FETA_OK;
  FETA_REQUIRE(y && fi, "unpack_heads: NULL pointer argument");
  unpack_heads_kernel<false><<<seg_grid((int64_t)H * N * dh), kSegThreads, 0, st>>>(y, fi, out, N, B, H, dh);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_unpack_heads_bwd(const float* d_out, const int64_t* fi, float* dy, int64_t N, int B, int nmax,
                                     int H, int dh, void* stream_) {
  (void)nmax;
  if (N == 0) return FETA_OK;
  FETA_REQUIRE(d_out && fi && dy, "unpack_heads_bwd: NULL pointer argument");
  unpack_heads_kernel<true><<<seg_grid((int64_t)H * N * dh), kSegThreads, 0, (cudaStream_t)stream_>>>(d_out, fi, dy, N,
                                                                                                      B, H, dh);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_segment_mean_fwd(const float* x, const int32_t* graph_ptr, float* out, int64_t G, int C,
                                     void* stream_) {
  if (G == 0) return FETA_OK;
  FETA_REQUIRE(x && graph_ptr && out && C >= 1, "segment_mean_fwd: bad argument");
  segment_mean_fwd_kernel<<<(unsigned)G, C < kSegThreads ? ((C + 31) / 32 * 32) : kSegThreads, 0,
                            (cudaStream_t)stream_>>>(x, graph_ptr, out, C);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_segment_mean_bwd(const float* d_out, const int32_t* graph_ptr, float* dx, int64_t G, int C,
                                     void* stream_) {
  if (G == 0) return FETA_OK;
  FETA_REQUIRE(d_out && graph_ptr && dx && C >= 1, "segment_mean_bwd: bad argument");
  segment_mean_bwd_kernel<<<(unsigned)G, C < kSegThreads ? ((C + 31) / 32 * 32) : kSegThreads, 0,
                            (cudaStream_t)stream_>>>(d_out, graph_ptr, dx, C);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_masked_mean_fwd(const float* x, int64_t sb, int64_t sn, const uint8_t* mask, float* out, int B,
                                    int nmax, int C, void* stream_) {
  if (B == 0) return FETA_OK;
  FETA_REQUIRE(x && mask && out && C >= 1 && nmax >= 1, "masked_mean_fwd: bad argument");
  masked_mean_fwd_kernel<<<(unsigned)B, kSegThreads, kSegThreads * sizeof(float), (cudaStream_t)stream_>>>(
      x, sb, sn, mask, out, nmax, C);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_masked_mean_bwd(const float* d_out, const uint8_t* mask, float* dx, int B, int nmax, int C,
                                    void* stream_) {
  if (B == 0) return FETA_OK;
  FETA_REQUIRE(d_out && mask && dx && C >= 1 && nmax >= 1, "masked_mean_bwd: bad argument");
  masked_mean_bwd_kernel<<<(unsigned)B, kSegThreads, 0, (cudaStream_t)stream_>>>(d_out, mask, dx, nmax, C);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_gather_rows(const float* padded, int64_t sb, int64_t sn, const int64_t* fi, float* packed, int64_t N,
                                int C, void* stream_) {
  if (N == 0) return FETA_OK;
  FETA_REQUIRE(padded && fi && packed && C >= 1, "gather_rows: bad argument");
  gather_rows_kernel<false><<<seg_grid(N * C), kSegThreads, 0, (cudaStream_t)stream_>>>(padded, fi, packed, sb, sn, N,
                                                                                        C);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_scatter_rows(const float* packed, const int64_t* fi, float* padded, int64_t sb, int64_t sn,
                                 int64_t N, int C, void* stream_) {
  if (N == 0) return FETA_OK;
  FETA_REQUIRE(padded && fi && packed && C >= 1, "scatter_rows: bad argument");
  gather_rows_kernel<true><<<seg_grid(N * C), kSegThreads, 0, (cudaStream_t)stream_>>>(packed, fi, padded, sb, sn, N,
                                                                                       C);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}
