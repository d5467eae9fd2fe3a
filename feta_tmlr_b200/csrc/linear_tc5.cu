// A6 (layer glue): the encoder layer's projections on the 5th-generation tensor cores (tcgen05 + TMEM),
// fp32-grade accuracy through the 3-term TF32 split -- the tcgen05 successor of dense_tc.cu (legacy mma.sync, which
// lost to the library's SIMT sgemm) and the replacement of that library sgemm on the default path.
//
//   Y[T, N]  = act(X[T, K] . W[N, K]^T + b)                       (nn.Linear forward)
//   dX[T, K] = (dY[T, N] . W[N, K]) * [mask > 0] + dres           (its input gradient, ReLU mask / residual fused)
//
// One CTA = 128 token rows x 64 output columns (grid.y walks the output columns: 111 CTAs for the ZINC in-projection
// instead of 37).  The reduction dimension is walked in chunks of 64: both operands of a chunk are written to shared
// memory in the UMMA canonical K-major layout as raw fp32 ("hi": the tensor core reads the top 19 bits) and the exact
// remainder ("lo"), one elected thread issues hi.hi + hi.lo + lo.hi as tcgen05.mma.kind::tf32 (M = 128, N = 64,
// K = 8), accumulators stay in TMEM across chunks, and the epilogue (tcgen05.ld, one thread per token row) applies
// bias / ReLU / mask / residual on the way out.
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace feta {
namespace lin5 {

using namespace tc;

constexpr int kThreads = 128, kBM = 128, kBN = 64, kKC = 64, kCores = kKC / 4;
constexpr uint32_t kABytes = kBM * kKC * 4, kBBytes = kBN * kKC * 4;      // one hi (or lo) slab

// Round-to-nearest split (unlike the Chebyshev tile kernel this slab is only ever read by the tensor core):
// hi = rna_tf32(x), lo = rna_tf32(x - hi): unbiased, ~2^-22 relative error per product instead of the ~2^-20
// one-sided error of the truncating split -- the model-level gradients sat at 1.2e-4 with the latter.
__device__ __forceinline__ void split4(float4 v, float4& hi, float4& lo) {
  hi = make_float4(tf32_hi(v.x), tf32_hi(v.y), tf32_hi(v.z), tf32_hi(v.w));
  lo = make_float4(tf32_hi(v.x - hi.x), tf32_hi(v.y - hi.y), tf32_hi(v.z - hi.z), tf32_hi(v.w - hi.w));
}

__device__ __forceinline__ uint32_t canon_off(int row, int q) {
  return (uint32_t)((((row >> 3) * kCores + q) << 7) + ((row & 7) << 4));
}

// The shared main loop: D[128, 64] (TMEM) = A[row0 .. row0+127, :] . B^T over K in chunks of 64.
// MODE 0: B[n][k] = W[n0 + n][k0 + k] (forward);  MODE 1: B[n][k] = W[k0 + k][n0 + n] (input gradient).
template <int MODE>
__device__ __forceinline__ void mainloop(const float* __restrict__ A, const float* __restrict__ W, int64_t T, int K,
                                         int ldw, int n0, unsigned char* smem, uint64_t* bar, uint32_t tmem) {
  unsigned char* aH = smem;
  unsigned char* aL = aH + kABytes;
  unsigned char* bH = aL + kABytes;
  unsigned char* bL = bH + kBBytes;
  const int tid = threadIdx.x;
  const int64_t row = (int64_t)blockIdx.x * kBM + tid;
  const bool live = row < T;
  const uint32_t idesc = make_idesc(kBM, kBN);
  const int nchunks = K / kKC;
  for (int c = 0; c < nchunks; ++c) {
    const int k0 = c * kKC;
    if (c > 0) mbar_wait_parity(smem_u32(bar), (uint32_t)((c - 1) & 1));   // the previous chunk's MMAs read smem
    // ---- A: this thread's token row, 16 x 16 bytes, issued together
    {
      float4 v[kCores];
      const float* src = A + row * K + k0;
#pragma unroll
      for (int q = 0; q < kCores; ++q)
        v[q] = live ? __ldg(reinterpret_cast<const float4*>(src + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < kCores; ++q) {
        const uint32_t off = canon_off(tid, q);
        float4 hi, lo;
        split4(v[q], hi, lo);
        *reinterpret_cast<float4*>(aH + off) = hi;
        *reinterpret_cast<float4*>(aL + off) = lo;
      }
    }
    // ---- B: 64 output columns x 64 reduction values of W; thread = (n = tid % 64, half of the cores)
    {
      const int n = tid & 63, qh = (tid >> 6) * (kCores / 2);
      float4 v[kCores / 2];
      if (MODE == 0) {
        const float* src = W + (int64_t)(n0 + n) * ldw + k0 + 4 * qh;
#pragma unroll
        for (int q = 0; q < kCores / 2; ++q) v[q] = __ldg(reinterpret_cast<const float4*>(src + 4 * q));
      } else {
        const float* src = W + (int64_t)(k0 + 4 * qh) * ldw + n0 + n;      // W[k][n]: coalesced across n
#pragma unroll
        for (int q = 0; q < kCores / 2; ++q)
          v[q] = make_float4(__ldg(src + (int64_t)(4 * q) * ldw), __ldg(src + (int64_t)(4 * q + 1) * ldw),
                             __ldg(src + (int64_t)(4 * q + 2) * ldw), __ldg(src + (int64_t)(4 * q + 3) * ldw));
      }
#pragma unroll
      for (int q = 0; q < kCores / 2; ++q) {
        const uint32_t off = canon_off(n, qh + q);
        float4 hi, lo;
        split4(v[q], hi, lo);
        *reinterpret_cast<float4*>(bH + off) = hi;
        *reinterpret_cast<float4*>(bL + off) = lo;
      }
    }
    fence_proxy_async();
    fence_before();
    __syncthreads();
    fence_after();
    if (tid == 0) {
      constexpr uint32_t sbo = kCores * 128;
#pragma unroll
      for (int ks = 0; ks < kKC / 8; ++ks) {
        const uint64_t dAh = make_desc(smem_u32(aH) + ks * 256, 128, sbo), dAl = make_desc(smem_u32(aL) + ks * 256, 128, sbo);
        const uint64_t dBh = make_desc(smem_u32(bH) + ks * 256, 128, sbo), dBl = make_desc(smem_u32(bL) + ks * 256, 128, sbo);
        mma_ss(tmem, dAh, dBh, idesc, (c > 0 || ks > 0) ? 1u : 0u);
        mma_ss(tmem, dAh, dBl, idesc, 1u);
        mma_ss(tmem, dAl, dBh, idesc, 1u);
      }
      mma_commit(smem_u32(bar));
    }
  }
  mbar_wait_parity(smem_u32(bar), (uint32_t)((nchunks - 1) & 1));
  fence_after();
}

__device__ __forceinline__ uint32_t prologue(uint64_t* bar, uint32_t* slot) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(smem_u32(slot), 64);
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    mbar_fence_init();
  }
  fence_before();
  __syncthreads();
  fence_after();
  return *slot;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) linear_tc5_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                                const float* __restrict__ bias,
                                                                const float* __restrict__ dres,
                                                                const float* __restrict__ mask_src, float* __restrict__ Y,
                                                                int64_t T, int K, int N, int ldw, int relu) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t row = (int64_t)blockIdx.x * kBM + tid;
  const int n0 = blockIdx.y * kBN;
  const bool live = row < T;
  const uint32_t tmem = prologue(&s_bar, &s_tmem);
  mainloop<MODE>(A, W, T, K, ldw, n0, smem, &s_bar, tmem);
  // ---- epilogue: one thread per token row
  const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
#pragma unroll
  for (int c0 = 0; c0 < kBN; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + lane_base + c0, v);
    if (live) {
      const int col = n0 + c0;
      float* dst = Y + row * N + col;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float4 o = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        if (MODE == 0) {
          if (bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + 4 * g));
            o.x += b.x, o.y += b.y, o.z += b.z, o.w += b.w;
          }
          if (relu) o.x = fmaxf(o.x, 0.f), o.y = fmaxf(o.y, 0.f), o.z = fmaxf(o.z, 0.f), o.w = fmaxf(o.w, 0.f);
        } else {
          if (mask_src) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(mask_src + row * N + col + 4 * g));
            o.x = m.x > 0.f ? o.x : 0.f, o.y = m.y > 0.f ? o.y : 0.f;
            o.z = m.z > 0.f ? o.z : 0.f, o.w = m.w > 0.f ? o.w : 0.f;
          }
          if (dres) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(dres + row * N + col + 4 * g));
            o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
          }
        }
        *reinterpret_cast<float4*>(dst + 4 * g) = o;
      }
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// Linear + residual + LayerNorm in one launch (N = 64 = d_model: a token's whole output row is in one thread):
//   z = res + bscale[row] * (X . W^T + b);  y = LayerNorm(z) * gamma + beta;  z, mean, rstd kept for the backward pass
__global__ void __launch_bounds__(kThreads, 2) linear_ln_tc5_kernel(
    const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
    const float* __restrict__ res, const float* __restrict__ bscale, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ z, float* __restrict__ mean,
    float* __restrict__ rstd, int64_t T, int K, float eps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ float s_g[kBN], s_b[kBN], s_bias[kBN];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t row = (int64_t)blockIdx.x * kBM + tid;
  const bool live = row < T;
  if (tid < kBN) {
    s_g[tid] = __ldg(gamma + tid);
    s_b[tid] = __ldg(beta + tid);
    s_bias[tid] = bias ? __ldg(bias + tid) : 0.0f;
  }
  const uint32_t tmem = prologue(&s_bar, &s_tmem);
  mainloop<0>(X, W, T, K, K, 0, smem, &s_bar, tmem);
  const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
  float v[kBN];
#pragma unroll
  for (int c0 = 0; c0 < kBN; c0 += 16) {
    float t16[16];
    tmem_ld16(tmem + lane_base + c0, t16);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[c0 + i] = t16[i];
  }
  if (live) {
    const float bs = bscale ? __ldg(bscale + row) : 1.0f;
    float sum = 0.0f;
#pragma unroll
    for (int g = 0; g < kBN / 4; ++g) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(res + row * kBN + 4 * g));
      v[4 * g] = fmaf(bs, v[4 * g] + s_bias[4 * g], r.x);
      v[4 * g + 1] = fmaf(bs, v[4 * g + 1] + s_bias[4 * g + 1], r.y);
      v[4 * g + 2] = fmaf(bs, v[4 * g + 2] + s_bias[4 * g + 2], r.z);
      v[4 * g + 3] = fmaf(bs, v[4 * g + 3] + s_bias[4 * g + 3], r.w);
      sum += (v[4 * g] + v[4 * g + 1]) + (v[4 * g + 2] + v[4 * g + 3]);
    }
    const float mu = sum * (1.0f / kBN);
    float q = 0.0f;
#pragma unroll
    for (int i = 0; i < kBN; ++i) q = fmaf(v[i] - mu, v[i] - mu, q);
    const float rs = 1.0f / sqrtf(q * (1.0f / kBN) + eps);
#pragma unroll
    for (int g = 0; g < kBN / 4; ++g) {
      *reinterpret_cast<float4*>(z + row * kBN + 4 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
      float4 o;
      o.x = (v[4 * g] - mu) * rs * s_g[4 * g] + s_b[4 * g];
      o.y = (v[4 * g + 1] - mu) * rs * s_g[4 * g + 1] + s_b[4 * g + 1];
      o.z = (v[4 * g + 2] - mu) * rs * s_g[4 * g + 2] + s_b[4 * g + 2];
      o.w = (v[4 * g + 3] - mu) * rs * s_g[4 * g + 3] + s_b[4 * g + 3];
      *reinterpret_cast<float4*>(y + row * kBN + 4 * g) = o;
    }
    mean[row] = mu;
    rstd[row] = rs;
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

// =====================================================================================================
// The layer's tail in ONE launch (d_model = 64, dim_feedforward = 128 -- every BASELINE config):
//     z1 = res + bscale * (o . Wo^T + bo);   y1 = LN1(z1)            (out-projection, degree scale, residual, norm1)
//     h  = relu(y1 . W1^T + b1)                                        (linear1)
//     z2 = y1 + h . W2^T + b2;               y2 = LN2(z2)            (linear2, residual, norm2)
// Three chained tcgen05 GEMMs per 128-token tile: the activations never leave the SM between them -- y1 and h go
// from the epilogue's registers straight back into shared memory as the next A operand (hi / lo, canonical layout)
// -- and everything the backward pass needs (z1, mean1, rstd1, y1, h, z2, mean2, rstd2) is written once.
// Replaces 5 launches (out_proj GEMM, add+LayerNorm, linear1 GEMM, linear2 GEMM, add+LayerNorm) of the chain.
// =====================================================================================================
constexpr int kD = 64, kDFF = 128;
constexpr uint32_t kChunkA = 2 * kABytes;                 // hi + lo of one 128 x 64 A chunk
constexpr size_t kTailSmem = 2 * (size_t)kChunkA + 4 * (size_t)kBBytes;   // A: two chunks; B: up to 128 x 64 hi + lo

// stage W[rows n0..n0+NR-1][k0..k0+63] (row-major, leading dimension ldw) as a canonical hi / lo B slab
template <int NR>
__device__ __forceinline__ void stage_w(unsigned char* bH, unsigned char* bL, const float* __restrict__ W, int ldw,
                                        int k0) {
  constexpr int PER = NR * kCores / kThreads;             // 16-byte cores per thread
  const int tid = threadIdx.x;
  float4 v[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int idx = tid + i * kThreads, n = idx / kCores, q = idx % kCores;
    v[i] = __ldg(reinterpret_cast<const float4*>(W + (int64_t)n * ldw + k0 + 4 * q));
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int idx = tid + i * kThreads, n = idx / kCores, q = idx % kCores;
    float4 hi, lo;
    split4(v[i], hi, lo);
    const uint32_t off = canon_off(n, q);
    *reinterpret_cast<float4*>(bH + off) = hi;
    *reinterpret_cast<float4*>(bL + off) = lo;
  }
}

// D[128, N] (+)= A_chunk . B^T over one 64-wide K chunk
template <int N>
__device__ __forceinline__ void issue_chunk(uint32_t tmem_d, const unsigned char* aH, const unsigned char* aL,
                                            const unsigned char* bH, const unsigned char* bL, bool accumulate) {
  const uint32_t idesc = make_idesc(kBM, N);
  constexpr uint32_t sbo = kCores * 128;
#pragma unroll
  for (int ks = 0; ks < kKC / 8; ++ks) {
    const uint64_t dAh = make_desc(smem_u32(aH) + ks * 256, 128, sbo), dAl = make_desc(smem_u32(aL) + ks * 256, 128, sbo);
    const uint64_t dBh = make_desc(smem_u32(bH) + ks * 256, 128, sbo), dBl = make_desc(smem_u32(bL) + ks * 256, 128, sbo);
    mma_ss(tmem_d, dAh, dBh, idesc, (accumulate || ks > 0) ? 1u : 0u);
    mma_ss(tmem_d, dAh, dBl, idesc, 1u);
    mma_ss(tmem_d, dAl, dBh, idesc, 1u);
  }
}

__device__ __forceinline__ void sync_for_mma() {
  fence_proxy_async();
  fence_before();
  __syncthreads();
  fence_after();
}

// LayerNorm of the 64 values a thread holds; returns mean / rstd
__device__ __forceinline__ void ln64(const float (&v)[kD], float eps, float& mu, float& rs) {
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < kD; i += 4) s += (v[i] + v[i + 1]) + (v[i + 2] + v[i + 3]);
  mu = s * (1.0f / kD);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < kD; ++i) q = fmaf(v[i] - mu, v[i] - mu, q);
  rs = 1.0f / sqrtf(q * (1.0f / kD) + eps);
}

__global__ void __launch_bounds__(kThreads, 1) layer_tail_fwd_kernel(
    const float* __restrict__ o, const float* __restrict__ res, const float* __restrict__ bscale,
    const float* __restrict__ Wo, const float* __restrict__ bo, const float* __restrict__ g1,
    const float* __restrict__ be1, const float* __restrict__ W1, const float* __restrict__ b1,
    const float* __restrict__ W2, const float* __restrict__ b2, const float* __restrict__ g2,
    const float* __restrict__ be2, float* __restrict__ z1, float* __restrict__ mean1, float* __restrict__ rstd1,
    float* __restrict__ y1, float* __restrict__ h, float* __restrict__ z2, float* __restrict__ mean2,
    float* __restrict__ rstd2, float* __restrict__ y2, int64_t T, float eps1, float eps2) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ float s_par[6 * kD + kDFF];                  // bo, g1, be1, b2, g2, be2 [64 each], b1 [128]
  unsigned char* A0h = smem;                              // A chunk 0: hi, lo;  chunk 1: hi, lo
  unsigned char* A0l = A0h + kABytes;
  unsigned char* A1h = A0l + kABytes;
  unsigned char* A1l = A1h + kABytes;
  unsigned char* Bh = A1l + kABytes;                      // B: up to 128 rows x 64 k (hi 32 KB, lo 32 KB)
  unsigned char* Bl = Bh + 2 * kBBytes;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t row = (int64_t)blockIdx.x * kBM + tid;
  const bool live = row < T;
  if (tid < kD) {
    s_par[tid] = bo ? __ldg(bo + tid) : 0.0f;
    s_par[kD + tid] = __ldg(g1 + tid);
    s_par[2 * kD + tid] = __ldg(be1 + tid);
    s_par[3 * kD + tid] = b2 ? __ldg(b2 + tid) : 0.0f;
    s_par[4 * kD + tid] = __ldg(g2 + tid);
    s_par[5 * kD + tid] = __ldg(be2 + tid);
  }
  s_par[6 * kD + tid] = b1 ? __ldg(b1 + tid) : 0.0f;     // kThreads == kDFF
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 256);
  if (tid == 0) {
    mbar_init(smem_u32(&s_bar), 1);
    mbar_fence_init();
  }
  // ---- stage 1 operands: A = o rows, B = Wo
  {
    float4 v[kCores];
    const float* src = o + row * kD;
#pragma unroll
    for (int q = 0; q < kCores; ++q)
      v[q] = live ? __ldg(reinterpret_cast<const float4*>(src + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < kCores; ++q) {
      float4 hi, lo;
      split4(v[q], hi, lo);
      *reinterpret_cast<float4*>(A0h + canon_off(tid, q)) = hi;
      *reinterpret_cast<float4*>(A0l + canon_off(tid, q)) = lo;
    }
  }
  stage_w<kD>(Bh, Bl, Wo, kD, 0);
  sync_for_mma();
  const uint32_t tmem = s_tmem;
  const uint32_t D1 = tmem, D2 = tmem + 64, D3 = tmem + 192;
  const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
  if (tid == 0) {
    issue_chunk<kD>(D1, A0h, A0l, Bh, Bl, false);
    mma_commit(smem_u32(&s_bar));
  }
  mbar_wait_parity(smem_u32(&s_bar), 0);
  fence_after();
  // ---- epilogue 1: z1, LN1 -> y1 (registers + global + next A operand); B <- W1
  float yv[kD];
#pragma unroll
  for (int c0 = 0; c0 < kD; c0 += 16) {
    float t16[16];
    tmem_ld16(D1 + lane_base + c0, t16);
#pragma unroll
    for (int i = 0; i < 16; ++i) yv[c0 + i] = t16[i];
  }
  {
    const float bs = (bscale && live) ? __ldg(bscale + row) : 1.0f;
#pragma unroll
    for (int g = 0; g < kD / 4; ++g) {
      const float4 r = live ? __ldg(reinterpret_cast<const float4*>(res + row * kD + 4 * g)) : make_float4(0.f, 0.f, 0.f, 0.f);
      yv[4 * g] = fmaf(bs, yv[4 * g] + s_par[4 * g], r.x);
      yv[4 * g + 1] = fmaf(bs, yv[4 * g + 1] + s_par[4 * g + 1], r.y);
      yv[4 * g + 2] = fmaf(bs, yv[4 * g + 2] + s_par[4 * g + 2], r.z);
      yv[4 * g + 3] = fmaf(bs, yv[4 * g + 3] + s_par[4 * g + 3], r.w);
    }
    float mu, rs;
    ln64(yv, eps1, mu, rs);
#pragma unroll
    for (int g = 0; g < kD / 4; ++g) {
      if (live) *reinterpret_cast<float4*>(z1 + row * kD + 4 * g) = make_float4(yv[4 * g], yv[4 * g + 1], yv[4 * g + 2], yv[4 * g + 3]);
#pragma unroll
      for (int i = 0; i < 4; ++i) yv[4 * g + i] = (yv[4 * g + i] - mu) * rs * s_par[kD + 4 * g + i] + s_par[2 * kD + 4 * g + i];
      const float4 yq = make_float4(yv[4 * g], yv[4 * g + 1], yv[4 * g + 2], yv[4 * g + 3]);
      if (live) *reinterpret_cast<float4*>(y1 + row * kD + 4 * g) = yq;
      float4 hi, lo;
      split4(yq, hi, lo);
      *reinterpret_cast<float4*>(A0h + canon_off(tid, g)) = hi;
      *reinterpret_cast<float4*>(A0l + canon_off(tid, g)) = lo;
    }
    if (live) {
      mean1[row] = mu;
      rstd1[row] = rs;
    }
  }
  stage_w<kDFF>(Bh, Bl, W1, kD, 0);
  fence_before();                       // the tcgen05.ld of D1 are ordered before the next MMAs
  sync_for_mma();
  if (tid == 0) {
    issue_chunk<kDFF>(D2, A0h, A0l, Bh, Bl, false);
    mma_commit(smem_u32(&s_bar));
  }
  mbar_wait_parity(smem_u32(&s_bar), 1);
  fence_after();
  // ---- epilogue 2: h = relu(. + b1) -> global + next A operand (two 64-wide chunks); B <- W2 (two K chunks)
#pragma unroll
  for (int c0 = 0; c0 < kDFF; c0 += 16) {
    float t16[16];
    tmem_ld16(D2 + lane_base + c0, t16);
    unsigned char* ah = (c0 < 64) ? A0h : A1h;
    unsigned char* al = (c0 < 64) ? A0l : A1l;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float4 hq;
      hq.x = fmaxf(t16[4 * g] + s_par[6 * kD + c0 + 4 * g], 0.f);
      hq.y = fmaxf(t16[4 * g + 1] + s_par[6 * kD + c0 + 4 * g + 1], 0.f);
      hq.z = fmaxf(t16[4 * g + 2] + s_par[6 * kD + c0 + 4 * g + 2], 0.f);
      hq.w = fmaxf(t16[4 * g + 3] + s_par[6 * kD + c0 + 4 * g + 3], 0.f);
      if (live) *reinterpret_cast<float4*>(h + row * kDFF + c0 + 4 * g) = hq;
      float4 hi, lo;
      split4(hq, hi, lo);
      const uint32_t off = canon_off(tid, ((c0 & 63) >> 2) + g);
      *reinterpret_cast<float4*>(ah + off) = hi;
      *reinterpret_cast<float4*>(al + off) = lo;
    }
  }
  stage_w<kD>(Bh, Bl, W2, kDFF, 0);                               // K chunk 0: W2[:, 0..63]
  stage_w<kD>(Bh + kBBytes, Bl + kBBytes, W2, kDFF, kKC);         // K chunk 1: W2[:, 64..127]
  fence_before();
  sync_for_mma();
  if (tid == 0) {
    issue_chunk<kD>(D3, A0h, A0l, Bh, Bl, false);
    issue_chunk<kD>(D3, A1h, A1l, Bh + kBBytes, Bl + kBBytes, true);
    mma_commit(smem_u32(&s_bar));
  }
  mbar_wait_parity(smem_u32(&s_bar), 0);
  fence_after();
  // ---- epilogue 3: z2 = y1 + . + b2, LN2 -> y2
  {
    float zv[kD];
#pragma unroll
    for (int c0 = 0; c0 < kD; c0 += 16) {
      float t16[16];
      tmem_ld16(D3 + lane_base + c0, t16);
#pragma unroll
      for (int i = 0; i < 16; ++i) zv[c0 + i] = yv[c0 + i] + t16[i] + s_par[3 * kD + c0 + i];
    }
    float mu, rs;
    ln64(zv, eps2, mu, rs);
    if (live) {
#pragma unroll
      for (int g = 0; g < kD / 4; ++g) {
        *reinterpret_cast<float4*>(z2 + row * kD + 4 * g) = make_float4(zv[4 * g], zv[4 * g + 1], zv[4 * g + 2], zv[4 * g + 3]);
        float4 yq;
        yq.x = (zv[4 * g] - mu) * rs * s_par[4 * kD + 4 * g] + s_par[5 * kD + 4 * g];
        yq.y = (zv[4 * g + 1] - mu) * rs * s_par[4 * kD + 4 * g + 1] + s_par[5 * kD + 4 * g + 1];
        yq.z = (zv[4 * g + 2] - mu) * rs * s_par[4 * kD + 4 * g + 2] + s_par[5 * kD + 4 * g + 2];
        yq.w = (zv[4 * g + 3] - mu) * rs * s_par[4 * kD + 4 * g + 3] + s_par[5 * kD + 4 * g + 3];
        *reinterpret_cast<float4*>(y2 + row * kD + 4 * g) = yq;
      }
      mean2[row] = mu;
      rstd2[row] = rs;
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

static bool eligible(int64_t T, int K, int N, const void* a, const void* w, const void* y, const void* p1,
                     const void* p2) {
  if (getenv("FETA_LINEAR_NO_TC5") != nullptr) return false;
  const uintptr_t ptrs = (uintptr_t)a | (uintptr_t)w | (uintptr_t)y | (uintptr_t)p1 | (uintptr_t)p2;
  return T >= 1 && K >= kKC && K % kKC == 0 && N >= kBN && N % kBN == 0 && K <= 1024 && N <= 1024 && (ptrs % 16) == 0;
}

template <int MODE>
static int launch(const float* A, const float* W, const float* bias, const float* dres, const float* mask_src, float* Y,
                  int64_t T, int K, int N, int ldw, int relu, cudaStream_t st) {
  const size_t smem = 2 * (size_t)kABytes + 2 * (size_t)kBBytes;
  FETA_CUDA(cudaFuncSetAttribute(linear_tc5_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(T, kBM), (unsigned)(N / kBN));
  linear_tc5_kernel<MODE><<<grid, kThreads, smem, st>>>(A, W, bias, dres, mask_src, Y, T, K, N, ldw, relu);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

static int launch_ln(const float* X, const float* W, const float* bias, const float* res, const float* bscale,
                     const float* gamma, const float* beta, float* y, float* z, float* mean, float* rstd, int64_t T,
                     int K, float eps, cudaStream_t st) {
  const size_t smem = 2 * (size_t)kABytes + 2 * (size_t)kBBytes;
  FETA_CUDA(cudaFuncSetAttribute(linear_ln_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  linear_ln_tc5_kernel<<<(unsigned)ceil_div(T, kBM), kThreads, smem, st>>>(X, W, bias, res, bscale, gamma, beta, y, z,
                                                                         mean, rstd, T, K, eps);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

}  // namespace lin5

// Y[T, out] = act(X[T, in] . W[out, in]^T + b); returns 1 when the shape is not eligible
int linear5_fwd_try(const float* X, const float* W, const float* bias, float* Y, int64_t T, int in, int out, int relu,
                    cudaStream_t st) {
  if (!lin5::eligible(T, in, out, X, W, Y, bias, nullptr)) return 1;
  return lin5::launch<0>(X, W, bias, nullptr, nullptr, Y, T, in, out, in, relu, st);
}

// dX[T, in] = (dY[T, out] . W[out, in]) * [mask > 0] + dres
int linear5_dx_try(const float* dY, const float* W, const float* dres, const float* mask_src, float* dX, int64_t T,
                   int in, int out, cudaStream_t st) {
  if (!lin5::eligible(T, out, in, dY, W, dX, dres, mask_src)) return 1;
  return lin5::launch<1>(dY, W, nullptr, dres, mask_src, dX, T, out, in, in, 0, st);
}

}  // namespace feta

// the whole tail of an encoder layer (d_model 64, dim_feedforward 128) in one launch (see include/feta_b200.h)
extern "C" int feta_layer_tail_supported(int d_model, int dff) {
  return d_model == 64 && dff == 128 && getenv("FETA_LINEAR_NO_TC5") == nullptr;
}

extern "C" int feta_layer_tail_fwd(const float* o, const float* res, const float* bscale, const float* Wo, const float* bo,
                                   const float* g1, const float* be1, const float* W1, const float* b1, const float* W2,
                                   const float* b2, const float* g2, const float* be2, float* z1, float* mean1,
                                   float* rstd1, float* y1, float* h, float* z2, float* mean2, float* rstd2, float* y2,
                                   int64_t T, int d_model, int dff, float eps1, float eps2, void* stream_) {
  using namespace feta;
  FETA_REQUIRE(T >= 0 && feta_layer_tail_supported(d_model, dff), "layer_tail_fwd: needs d_model 64 / dim_feedforward 128");
  if (T == 0) return FETA_OK;
  FETA_REQUIRE(o && res && Wo && g1 && be1 && W1 && W2 && g2 && be2 && z1 && mean1 && rstd1 && y1 && h && z2 && mean2 &&
                   rstd2 && y2, "layer_tail_fwd: NULL pointer");
  const uintptr_t ptrs = (uintptr_t)o | (uintptr_t)res | (uintptr_t)Wo | (uintptr_t)W1 | (uintptr_t)W2 | (uintptr_t)z1 |
                         (uintptr_t)y1 | (uintptr_t)h | (uintptr_t)z2 | (uintptr_t)y2;
  FETA_REQUIRE((ptrs % 16) == 0, "layer_tail_fwd: pointers must be 16-byte aligned");
  FETA_CUDA(cudaFuncSetAttribute(lin5::layer_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)lin5::kTailSmem));
  lin5::layer_tail_fwd_kernel<<<(unsigned)ceil_div(T, lin5::kBM), lin5::kThreads, lin5::kTailSmem,
                                (cudaStream_t)stream_>>>(o, res, bscale, Wo, bo, g1, be1, W1, b1, W2, b2, g2, be2, z1,
                                                         mean1, rstd1, y1, h, z2, mean2, rstd2, y2, T, eps1, eps2);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

// y = LayerNorm(res + bscale * (X . W^T + b)) * gamma + beta, out features = 64 (see include/feta_b200.h)
extern "C" int feta_linear_layernorm_supported(int in, int out) {
  return out == 64 && in >= 64 && in % 64 == 0 && in <= 1024 && getenv("FETA_LINEAR_NO_TC5") == nullptr;
}

namespace feta {
int linear_simt_ln_try(const float* X, const float* W, const float* bias, const float* res, const float* bscale,
                       const float* gamma, const float* beta, float* y, float* z, float* mean, float* rstd, int64_t T,
                       int in, int out, float eps, cudaStream_t st);
}
extern "C" int feta_linear_layernorm_simt_supported(int in, int out);

// impl: FETA_LINEAR_AUTO (SIMT kernel when eligible, else tcgen05), FETA_LINEAR_SIMT, FETA_LINEAR_TC5
extern "C" int feta_linear_layernorm_fwd_ex(const float* X, const float* W, const float* bias, const float* res,
                                            const float* bscale, const float* gamma, const float* beta, float* y,
                                            float* z, float* mean, float* rstd, int64_t T, int in, int out, float eps,
                                            int impl, void* stream_) {
  using namespace feta;
  const bool simt_ok = feta_linear_layernorm_simt_supported(in, out) != 0;
  const bool tc5_ok = feta_linear_layernorm_supported(in, out) != 0;
  FETA_REQUIRE(T >= 0 && (simt_ok || tc5_ok), "linear_layernorm_fwd: unsupported in=%d out=%d", in, out);
  if (T == 0) return FETA_OK;
  FETA_REQUIRE(X && W && res && gamma && beta && y && z && mean && rstd, "linear_layernorm_fwd: NULL pointer");
  const uintptr_t ptrs = (uintptr_t)X | (uintptr_t)W | (uintptr_t)res | (uintptr_t)y | (uintptr_t)z | (uintptr_t)bias;
  FETA_REQUIRE((ptrs % 16) == 0, "linear_layernorm_fwd: pointers must be 16-byte aligned");
  if (simt_ok && (impl == FETA_LINEAR_AUTO || impl == FETA_LINEAR_SIMT)) {
    const int rc = linear_simt_ln_try(X, W, bias, res, bscale, gamma, beta, y, z, mean, rstd, T, in, out, eps,
                                      (cudaStream_t)stream_);
    if (rc <= 0) return rc;
  }
  FETA_REQUIRE(tc5_ok, "linear_layernorm_fwd: in=%d out=%d not eligible for the tcgen05 kernel", in, out);
  return lin5::launch_ln(X, W, bias, res, bscale, gamma, beta, y, z, mean, rstd, T, in, eps, (cudaStream_t)stream_);
}

extern "C" int feta_linear_layernorm_fwd(const float* X, const float* W, const float* bias, const float* res,
                                         const float* bscale, const float* gamma, const float* beta, float* y, float* z,
                                         float* mean, float* rstd, int64_t T, int in, int out, float eps, void* stream_) {
  return feta_linear_layernorm_fwd_ex(X, W, bias, res, bscale, gamma, beta, y, z, mean, rstd, T, in, out, eps,
                                      FETA_LINEAR_AUTO, stream_);
}
