// A4: filter-coefficient path, collapsed (see include/feta_b200.h).
//
// Replaces DiffTransformerEncoderGenGCN.get_filter_coefficients (transformer/models.py:240-287):
// the reference builds, in a host Python loop, the complete directed graph on every (head, graph),
// uploads its [2, H*sum(n^2)] int64 edge list, and runs a PyG GCNConv over an all-ones feature
// matrix -- a [H*sum(n^2), ncoef] message tensor (15 GB at PATTERN shape).  Because the features
// are all ones the GCN output of node j is s_j * colsum(W) + b with the scalar s_j computed below
// straight from the attention tile; tanh + mean-pool (models.py:282-283) follow in one kernel.
#include <math.h>

#include "common.cuh"

namespace feta {

constexpr int kCoeffThreads = 256;

// one CTA per (graph b, head h)
__global__ void __launch_bounds__(1024) coeff_scalar_kernel(const float* __restrict__ attn,
                                                                    const uint8_t* __restrict__ mask,
                                                                    const int32_t* __restrict__ node_ptr,
                                                                    float* __restrict__ s, int H, int nmax,
                                                                    int64_t N) {
  extern __shared__ float smem[];
  float* dis = smem;                                   // [nmax] deg^-1/2 (0 for padding)
  float* loopw = dis + nmax;                           // [nmax]
  int* rank = reinterpret_cast<int*>(loopw + nmax);    // [nmax] packed index of a real position
  const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
  const uint8_t* mk = mask + (size_t)b * nmax;
  const float* a = attn + ((size_t)b * H + h) * nmax * nmax;

  // rank of every un-masked position (prefix count), by warp 0
  if (threadIdx.x < 32) {
    int run = 0;
    for (int j0 = 0; j0 < nmax; j0 += 32) {
      const int j = j0 + threadIdx.x;
      const bool real = (j < nmax) && (mk[j] == 0);
      const unsigned bal = __ballot_sync(0xffffffffu, real);
      if (j < nmax) rank[j] = real ? run + __popc(bal & ((1u << threadIdx.x) - 1u)) : -1;
      run += __popc(bal);
    }
  }
  __syncthreads();
  // thread (j, slice): column j (coalesced across j), rows slice, slice + nsl, ...; the slices of a column are folded
  // in shared memory in a fixed order.  Small graphs get several slices per column (ZINC: 4 x 64 threads), so the
  // per-thread chain of dependent loads is nmax / nsl long instead of nmax.
  float* red = reinterpret_cast<float*>(rank + nmax);  // [blockDim.x]
  const int jw = nmax >= (int)blockDim.x ? (int)blockDim.x : ((nmax + 31) & ~31);
  const int nsl = (int)blockDim.x / jw;
  const int jl = (int)threadIdx.x % jw, slice = (int)threadIdx.x / jw;
  // pass 1: deg_j = sum_{i != j} a_ij + loop_j  (gcn_norm: degree over the TARGET index)
  for (int j0 = 0; j0 < nmax; j0 += jw) {
    const int j = j0 + jl;
    const bool real = slice < nsl && j < nmax && rank[j] >= 0;
    float acc = 0.0f;
    if (real) {
#pragma unroll 4
      for (int i = slice; i < nmax; i += nsl)
        if (i != j && rank[i] >= 0) acc += a[(size_t)i * nmax + j];
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (slice == 0 && j < nmax) {
      float d = 0.0f, lw = 0.0f;
      if (real) {
        for (int k = 1; k < nsl; ++k) acc += red[k * jw + jl];
        const float ajj = a[(size_t)j * nmax + j];
        lw = (ajj != 0.0f) ? ajj : 1.0f;  // add_remaining_self_loops keeps an existing loop weight
        const float deg = acc + lw;
        d = deg > 0.0f ? 1.0f / sqrtf(deg) : 0.0f;
      }
      dis[j] = d;
      loopw[j] = lw;
    }
    __syncthreads();
  }
  // pass 2: s_j = dis_j * (sum_{i != j} dis_i a_ij + dis_j loop_j)
  const int64_t base = (int64_t)h * N + node_ptr[b];
  for (int j0 = 0; j0 < nmax; j0 += jw) {
    const int j = j0 + jl;
    const bool real = slice < nsl && j < nmax && rank[j] >= 0;
    float acc = 0.0f;
    if (real) {
#pragma unroll 4
      for (int i = slice; i < nmax; i += nsl)
        if (i != j && rank[i] >= 0) acc = fmaf(dis[i], a[(size_t)i * nmax + j], acc);
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    if (slice == 0 && real) {
      for (int k = 1; k < nsl; ++k) acc += red[k * jw + jl];
      s[base + rank[j]] = dis[j] * (acc + dis[j] * loopw[j]);
    }
    __syncthreads();
  }
}

// pooled[g, c] = mean_{j in g} tanh(s_j * wbar[c] + gbias[c]);  grid (G, ceil(C / 256))
__global__ void __launch_bounds__(kCoeffThreads) coeff_pool_fwd_kernel(const float* __restrict__ s,
                                                                      const int32_t* __restrict__ seg_lo,
                                                                      const int32_t* __restrict__ seg_hi,
                                                                      const float* __restrict__ wbar,
                                                                      const float* __restrict__ gbias,
                                                                      float* __restrict__ pooled, int C) {
  const int g = blockIdx.x;
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int lo = seg_lo[g], hi = seg_hi[g];
  const float w = wbar[c], bb = gbias[c];
  float acc = 0.0f;
  for (int j = lo; j < hi; ++j) acc += tanhf(fmaf(__ldg(s + j), w, bb));
  const int cnt = hi - lo;
  pooled[(size_t)g * C + c] = acc / (float)(cnt > 0 ? cnt : 1);  // scatter-mean clamps the count to 1
}

// partial[bx, 0, c] = sum_{g = bx mod nblk} sum_j dpool[g,c]/n_g * (1 - t^2) * s_j ; [bx, 1, c] without s_j
__global__ void __launch_bounds__(kCoeffThreads) coeff_pool_bwd_kernel(
    const float* __restrict__ s, const int32_t* __restrict__ seg_lo, const int32_t* __restrict__ seg_hi,
    const float* __restrict__ wbar,
    const float* __restrict__ gbias, const float* __restrict__ d_pooled, float* __restrict__ partial, int64_t G,
    int C) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float w = wbar[c], bb = gbias[c];
  float aw = 0.0f, ab = 0.0f;
  for (int64_t g = blockIdx.x; g < G; g += gridDim.x) {
    const int lo = seg_lo[g], hi = seg_hi[g];
    const int cnt = hi - lo;
    const float dg = d_pooled[(size_t)g * C + c] / (float)(cnt > 0 ? cnt : 1);
    float tw = 0.0f, tb = 0.0f;
    for (int j = lo; j < hi; ++j) {
      const float sj = __ldg(s + j);
      const float t = tanhf(fmaf(sj, w, bb));
      const float dt = 1.0f - t * t;
      tw = fmaf(dt, sj, tw);
      tb += dt;
    }
    aw = fmaf(dg, tw, aw);
    ab = fmaf(dg, tb, ab);
  }
  partial[((size_t)blockIdx.x * 2 + 0) * C + c] = aw;
  partial[((size_t)blockIdx.x * 2 + 1) * C + c] = ab;
}

// one warp per column: lanes walk the per-CTA partials, folded by shuffles (fixed order)
__global__ void __launch_bounds__(kCoeffThreads) coeff_pool_bwd_final_kernel(const float* __restrict__ partial,
                                                                            int nblk, int C,
                                                                            float* __restrict__ d_wbar,
                                                                            float* __restrict__ d_gbias) {
  const int c = blockIdx.x * (kCoeffThreads / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  float aw = 0.0f, ab = 0.0f;
  for (int b = lane; b < nblk; b += 32) {
    aw += partial[((size_t)b * 2 + 0) * C + c];
    ab += partial[((size_t)b * 2 + 1) * C + c];
  }
  aw = warp_sum(aw);
  ab = warp_sum(ab);
  if (lane == 0) {
    d_wbar[c] = aw;
    d_gbias[c] = ab;
  }
}

}  // namespace feta

using namespace feta;

extern "C" int feta_coeff_scalar(const float* attn, const uint8_t* mask, const int32_t* node_ptr, float* s, int B,
                                 int H, int nmax, int64_t N, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(B >= 0 && H >= 1 && nmax >= 0 && N >= 0, "coeff_scalar: bad sizes");
  if (B == 0 || nmax == 0 || N == 0) return FETA_OK;
  FETA_REQUIRE(attn && mask && node_ptr && s, "coeff_scalar: NULL pointer argument");
  // graphs wider than 64 nodes: 1024 threads, i.e. several row slices per column (PATTERN: 5 x 192) -- the kernel is
  // a chain of dependent loads per thread, nmax / slices long
  const int threads = nmax > 64 ? 1024 : kCoeffThreads;
  const size_t smem = ((size_t)nmax * 3 + threads) * sizeof(float);
  FETA_REQUIRE(smem <= 200 * 1024, "coeff_scalar: nmax=%d too large", nmax);
  if (smem > 48 * 1024)
    FETA_CUDA(cudaFuncSetAttribute(coeff_scalar_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  coeff_scalar_kernel<<<(unsigned)(B * H), threads, smem, st>>>(attn, mask, node_ptr, s, H, nmax, N);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_coeff_pool_fwd(const float* s, const int32_t* seg_lo, const int32_t* seg_hi, const float* wbar,
                                   const float* gbias,
                                   float* pooled, int64_t G, int C, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(G >= 0 && C >= 1, "coeff_pool_fwd: bad sizes");
  if (G == 0) return FETA_OK;
  FETA_REQUIRE(s && seg_lo && seg_hi && wbar && gbias && pooled, "coeff_pool_fwd: NULL pointer argument");
  dim3 grid((unsigned)G, (unsigned)ceil_div(C, kCoeffThreads));
  coeff_pool_fwd_kernel<<<grid, kCoeffThreads, 0, st>>>(s, seg_lo, seg_hi, wbar, gbias, pooled, C);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_coeff_pool_bwd(const float* s, const int32_t* seg_lo, const int32_t* seg_hi, const float* wbar,
                                   const float* gbias,
                                   const float* d_pooled, float* d_wbar, float* d_gbias, float* partial, int nblk,
                                   int64_t G, int C, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(G >= 0 && C >= 1 && nblk >= 1, "coeff_pool_bwd: bad sizes");
  FETA_REQUIRE(s && seg_lo && seg_hi && wbar && gbias && d_pooled && d_wbar && d_gbias && partial,
               "coeff_pool_bwd: NULL pointer argument");
  dim3 grid((unsigned)nblk, (unsigned)ceil_div(C, kCoeffThreads));
  coeff_pool_bwd_kernel<<<grid, kCoeffThreads, 0, st>>>(s, seg_lo, seg_hi, wbar, gbias, d_pooled, partial, G, C);
  FETA_LAUNCH_CHECK();
  coeff_pool_bwd_final_kernel<<<(unsigned)ceil_div(C, kCoeffThreads / 32), kCoeffThreads, 0, st>>>(partial, nblk, C,
                                                                                              d_wbar, d_gbias);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}
