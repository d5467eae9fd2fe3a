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

__device__ __forceinline__ uint32_t canon_off(int row, int q) {
  return (uint32_t)((((row >> 3) * kCores + q) << 7) + ((row & 7) << 4));
}

// MODE 0: forward (B[n][k] = W[n0 + n][k0 + k]);  MODE 1: input gradient (B[n][k] = W[k0 + k][n0 + n])
template <int MODE>
__global__ void __launch_bounds__(kThreads, 2) linear_tc5_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                                const float* __restrict__ bias,
                                                                const float* __restrict__ dres,
                                                                const float* __restrict__ mask_src, float* __restrict__ Y,
                                                                int64_t T, int K, int N, int ldw, int relu) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  unsigned char* aH = smem;
  unsigned char* aL = aH + kABytes;
  unsigned char* bH = aL + kABytes;
  unsigned char* bL = bH + kBBytes;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t row = (int64_t)blockIdx.x * kBM + tid;
  const int n0 = blockIdx.y * kBN;
  const bool live = row < T;

  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 64);
  if (tid == 0) {
    mbar_init(smem_u32(&s_bar), 1);
    mbar_fence_init();
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = make_idesc(kBM, kBN);
  const int nchunks = K / kKC;

  for (int c = 0; c < nchunks; ++c) {
    const int k0 = c * kKC;
    if (c > 0) mbar_wait_parity(smem_u32(&s_bar), (uint32_t)((c - 1) & 1));   // the previous chunk's MMAs read smem
    // ---- A: this thread's token row, 16 x 16 bytes, issued together
    {
      float4 v[kCores];
      const float* src = A + row * K + k0;
#pragma unroll
      for (int q = 0; q < kCores; ++q) v[q] = live ? __ldg(reinterpret_cast<const float4*>(src + 4 * q)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < kCores; ++q) {
        const uint32_t off = canon_off(tid, q);
        *reinterpret_cast<float4*>(aH + off) = v[q];
        *reinterpret_cast<float4*>(aL + off) = tf32_lo4(v[q]);
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
        *reinterpret_cast<float4*>(bH + off) = v[q];
        *reinterpret_cast<float4*>(bL + off) = tf32_lo4(v[q]);
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
      mma_commit(smem_u32(&s_bar));
    }
  }
  mbar_wait_parity(smem_u32(&s_bar), (uint32_t)((nchunks - 1) & 1));
  fence_after();
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
