// A6 (layer glue): the layer's small projections on the tensor cores with fp32-grade accuracy.
//
// The encoder layer's Linear layers are [T, 64..192] x [64..192] GEMMs (T = Nmax*B ~ 4-11k tokens): far too
// small to fill the GPU through the library's SIMT sgemm (6-8 us per launch inside the step graph, the
// largest block of the critical path after attention).  Here a CTA owns BM token rows x 64 output columns
// with the whole reduction dimension resident in shared memory (one load phase, no k-pipeline needed at
// K <= 256); each warp computes a 16 x 64 tile as m16n8k8 TF32 MMAs with BOTH operands split
// x = trunc19(x) + lo and three MMAs per product (hi.hi + hi.lo + lo.hi, the same scheme as the Chebyshev
// warp kernel), so results agree with fp32 to ~1e-6 relative -- the parity tests hold at 1e-4.
// Epilogues fuse what the reference's layer does around each GEMM: bias, ReLU, and in the backward the
// ReLU mask and the residual-gradient accumulation.
//
//   feta_linear_fwd:  Y[T,N]  = act(X[T,K] . W[N,K]^T + b)            (nn.Linear forward, W as PyTorch stores it)
//   feta_linear_dx:   dX[T,K] = (dY[T,N] . W[N,K]) * [mask > 0] + dres  (its input gradient)
#include <stdlib.h>

#include "common.cuh"

namespace feta {

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x);                                               // the MMA truncates to 19 bits itself
  lo = __float_as_uint(x - __uint_as_float(hi & 0xFFFFE000u));           // exact remainder
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int kBN = 64;   // output columns per CTA (8 n-tiles per warp)

// acc[nt] += A[16 x KR] . B  for one warp.  As: the warp's 16 rows, row stride lda (lda/4 odd => conflict-free).
// B_KMAJOR: Bs[n][k] with stride ldb (ldb/4 odd);  else Bs[k][n] with stride ldb (ldb % 32 == 8).
template <bool B_KMAJOR>
__device__ __forceinline__ void warp_gemm_16x64(float (&acc)[8][4], const float* __restrict__ As, int lda,
                                                const float* __restrict__ Bs, int ldb, int KR, int lane) {
  const int g = lane >> 2, tq = lane & 3;
  const float* a_lo = As + g * lda + tq;
  const float* a_hi = a_lo + 8 * lda;
  for (int ks = 0; ks < KR; ks += 8) {
    uint32_t ah[4], al[4];
    split_tf32(a_lo[ks], ah[0], al[0]);
    split_tf32(a_hi[ks], ah[1], al[1]);
    split_tf32(a_lo[ks + 4], ah[2], al[2]);
    split_tf32(a_hi[ks + 4], ah[3], al[3]);
    uint32_t bh[8][2], bl[8][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (B_KMAJOR) {
        const float* bp = Bs + (8 * nt + g) * ldb + ks + tq;
        split_tf32(bp[0], bh[nt][0], bl[nt][0]);
        split_tf32(bp[4], bh[nt][1], bl[nt][1]);
      } else {
        const float* bp = Bs + (ks + tq) * ldb + 8 * nt + g;
        split_tf32(bp[0], bh[nt][0], bl[nt][0]);
        split_tf32(bp[4 * ldb], bh[nt][1], bl[nt][1]);
      }
    }
    // the three split terms are issued nt-interleaved: consecutive MMAs hit different accumulators
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) mma_tf32(acc[nt], ah, bh[nt]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) mma_tf32(acc[nt], ah, bl[nt]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) mma_tf32(acc[nt], al, bh[nt]);
  }
}

// rows [r0, r0 + rows) x cols [c0, c0 + cols) of a row-major [R, C] matrix -> shared tile with stride ld;
// out-of-range entries are zero.  cols, c0 and C are multiples of 4, src 16-byte aligned.
__device__ __forceinline__ void load_tile(float* __restrict__ dst, int ld, const float* __restrict__ src, int64_t R,
                                          int C, int64_t r0, int rows, int c0, int cols) {
  const int q4 = cols >> 2;
  // (r, q) advance incrementally: one division per thread, not one per element
  int r = threadIdx.x / q4, q = threadIdx.x - r * q4;
  const int dr = blockDim.x / q4, dq = blockDim.x - dr * q4;
  // asynchronous 16-byte copies (LDGSTS): every copy of the tile is in flight at once instead of one L2 round
  // trip per loop iteration; out-of-range chunks are zero-filled (src-size 0)
  while (r < rows) {
    const int64_t row = r0 + r;
    const int col = c0 + 4 * q;
    const bool ok = row < R && col < C;
    const float* sp = ok ? src + row * C + col : src;
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst + r * ld + 4 * q);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(sp), "r"(ok ? 16 : 0) : "memory");
    r += dr, q += dq;
    if (q >= q4) q -= q4, ++r;
  }
}
__device__ __forceinline__ void tiles_ready() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
}

template <int BM>
__global__ void __launch_bounds__(BM * 2) linear_fwd_tc_kernel(const float* __restrict__ X, const float* __restrict__ W,
                                                              const float* __restrict__ bias, float* __restrict__ Y,
                                                              int64_t T, int K, int N, int relu) {
  extern __shared__ float4 smem_f4[];
  const int ld = K + 4;
  float* Xs = reinterpret_cast<float*>(smem_f4);   // [BM][ld]
  float* Ws = Xs + BM * ld;                        // [64][ld]   rows = output features (K-major B operand)
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int col0 = blockIdx.y * kBN;
  load_tile(Xs, ld, X, T, K, row0, BM, 0, K);
  load_tile(Ws, ld, W, N, K, col0, kBN, 0, K);
  tiles_ready();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
  warp_gemm_16x64<true>(acc, Xs + warp * 16 * ld, ld, Ws, ld, K, lane);
  const int64_t ra = row0 + warp * 16 + g, rb = ra + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = col0 + 8 * nt + 2 * tq;
    if (col >= N) continue;
    const float b0 = bias ? __ldg(bias + col) : 0.0f, b1 = bias ? __ldg(bias + col + 1) : 0.0f;
    float2 va = make_float2(acc[nt][0] + b0, acc[nt][1] + b1), vb = make_float2(acc[nt][2] + b0, acc[nt][3] + b1);
    if (relu) {
      va.x = fmaxf(va.x, 0.f), va.y = fmaxf(va.y, 0.f);
      vb.x = fmaxf(vb.x, 0.f), vb.y = fmaxf(vb.y, 0.f);
    }
    if (ra < T) *reinterpret_cast<float2*>(Y + ra * N + col) = va;
    if (rb < T) *reinterpret_cast<float2*>(Y + rb * N + col) = vb;
  }
}

template <int BM>
__global__ void __launch_bounds__(BM * 2) linear_dx_tc_kernel(const float* __restrict__ dY, const float* __restrict__ W,
                                                             const float* __restrict__ dres,
                                                             const float* __restrict__ mask_src,
                                                             float* __restrict__ dX, int64_t T, int K, int N) {
  extern __shared__ float4 smem_f4[];
  const int lda = N + 4;
  constexpr int ldb = kBN + 8;
  float* Ys = reinterpret_cast<float*>(smem_f4);   // [BM][lda]   dY rows (reduction over N)
  float* Ws = Ys + BM * lda;                       // [N][ldb]    W[:, col0 .. col0+63]
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int col0 = blockIdx.y * kBN;
  load_tile(Ys, lda, dY, T, N, row0, BM, 0, N);
  load_tile(Ws, ldb, W, N, K, 0, N, col0, kBN);
  tiles_ready();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.0f;
  warp_gemm_16x64<false>(acc, Ys + warp * 16 * lda, lda, Ws, ldb, N, lane);
  const int64_t ra = row0 + warp * 16 + g, rb = ra + 8;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = col0 + 8 * nt + 2 * tq;
    if (col >= K) continue;
    float2 va = make_float2(acc[nt][0], acc[nt][1]), vb = make_float2(acc[nt][2], acc[nt][3]);
    if (ra < T) {
      if (mask_src) {
        const float2 m = __ldg(reinterpret_cast<const float2*>(mask_src + ra * K + col));
        va.x = m.x > 0.f ? va.x : 0.f, va.y = m.y > 0.f ? va.y : 0.f;
      }
      if (dres) {
        const float2 r = __ldg(reinterpret_cast<const float2*>(dres + ra * K + col));
        va.x += r.x, va.y += r.y;
      }
      *reinterpret_cast<float2*>(dX + ra * K + col) = va;
    }
    if (rb < T) {
      if (mask_src) {
        const float2 m = __ldg(reinterpret_cast<const float2*>(mask_src + rb * K + col));
        vb.x = m.x > 0.f ? vb.x : 0.f, vb.y = m.y > 0.f ? vb.y : 0.f;
      }
      if (dres) {
        const float2 r = __ldg(reinterpret_cast<const float2*>(dres + rb * K + col));
        vb.x += r.x, vb.y += r.y;
      }
      *reinterpret_cast<float2*>(dX + rb * K + col) = vb;
    }
  }
}

static inline bool pick_bm64(int64_t T, int cols) {   // 64-row CTAs only when they still give >= 2 waves
  return ceil_div(T, 64) * ceil_div(cols, kBN) >= 2 * kNumSMs;
}

}  // namespace feta

using namespace feta;

namespace feta {
int linear5_fwd_try(const float* X, const float* W, const float* bias, float* Y, int64_t T, int in, int out, int relu,
                    cudaStream_t st);
int linear5_dx_try(const float* dY, const float* W, const float* dres, const float* mask_src, float* dX, int64_t T,
                   int in, int out, cudaStream_t st);
}

extern "C" int feta_linear_tc5_supported(int in, int out) {
  return in >= 64 && out >= 64 && in % 64 == 0 && out % 64 == 0 && in <= 1024 && out <= 1024 &&
         getenv("FETA_LINEAR_NO_TC5") == nullptr;
}

extern "C" int feta_linear_tc_supported(int in, int out) {
  return in >= 8 && out >= 8 && in % 8 == 0 && out % 8 == 0 && in <= 256 && out <= 256;
}

namespace feta {
int linear_simt_fwd_try(const float* X, const float* W, const float* bias, float* Y, int64_t T, int in, int out, int relu,
                        cudaStream_t st);
int linear_simt_dx_try(const float* dY, const float* W, const float* dres, const float* mask_src, float* dX, int64_t T,
                       int in, int out, cudaStream_t st);
}

extern "C" int feta_linear_fwd(const float* X, const float* W, const float* bias, float* Y, int64_t T, int in, int out,
                               int relu, void* stream_) {
  return feta_linear_fwd_ex(X, W, bias, Y, T, in, out, relu, FETA_LINEAR_AUTO, stream_);
}

extern "C" int feta_linear_dx(const float* dY, const float* W, const float* dres, const float* mask_src, float* dX,
                              int64_t T, int in, int out, void* stream_) {
  return feta_linear_dx_ex(dY, W, dres, mask_src, dX, T, in, out, FETA_LINEAR_AUTO, stream_);
}

extern "C" int feta_linear_fwd_ex(const float* X, const float* W, const float* bias, float* Y, int64_t T, int in, int out,
                                  int relu, int impl, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(T >= 0 && (feta_linear_tc_supported(in, out) || feta_linear_tc5_supported(in, out)),
               "linear_fwd: unsupported shape T=%lld in=%d out=%d",
               (long long)T, in, out);
  if (T == 0) return FETA_OK;
  FETA_REQUIRE(X && W && Y, "linear_fwd: NULL pointer argument");
  FETA_REQUIRE(((uintptr_t)X % 16) == 0 && ((uintptr_t)W % 16) == 0 && ((uintptr_t)Y % 8) == 0,
               "linear_fwd: X/W must be 16-byte aligned");
  if (impl == FETA_LINEAR_AUTO || impl == FETA_LINEAR_SIMT) {   // fp32 CUDA-core latency kernel (linear_simt.cu)
    const int rc = linear_simt_fwd_try(X, W, bias, Y, T, in, out, relu, st);
    if (rc <= 0) return rc;
  }
  if (impl != FETA_LINEAR_MMA) {   // tcgen05 path (linear_tc5.cu) when the shape is a multiple of its 128 x 64 x 64 tiles
    const int rc = linear5_fwd_try(X, W, bias, Y, T, in, out, relu, st);
    if (rc <= 0) return rc;
  }
  const int ld = in + 4;
  if (pick_bm64(T, out)) {
    const size_t smem = (size_t)(64 + kBN) * ld * sizeof(float);
    FETA_CUDA(cudaFuncSetAttribute(linear_fwd_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(T, 64), (unsigned)ceil_div(out, kBN));
    linear_fwd_tc_kernel<64><<<grid, 128, smem, st>>>(X, W, bias, Y, T, in, out, relu);
  } else {
    const size_t smem = (size_t)(32 + kBN) * ld * sizeof(float);
    FETA_CUDA(cudaFuncSetAttribute(linear_fwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(T, 32), (unsigned)ceil_div(out, kBN));
    linear_fwd_tc_kernel<32><<<grid, 64, smem, st>>>(X, W, bias, Y, T, in, out, relu);
  }
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_linear_dx_ex(const float* dY, const float* W, const float* dres, const float* mask_src, float* dX,
                                 int64_t T, int in, int out, int impl, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(T >= 0 && (feta_linear_tc_supported(in, out) || feta_linear_tc5_supported(in, out)),
               "linear_dx: unsupported shape T=%lld in=%d out=%d",
               (long long)T, in, out);
  if (T == 0) return FETA_OK;
  FETA_REQUIRE(dY && W && dX, "linear_dx: NULL pointer argument");
  FETA_REQUIRE(((uintptr_t)dY % 16) == 0 && ((uintptr_t)W % 16) == 0 && ((uintptr_t)dX % 8) == 0 &&
                   ((uintptr_t)dres % 8) == 0 && ((uintptr_t)mask_src % 8) == 0,
               "linear_dx: pointers must be 16-byte (dY, W) / 8-byte aligned");
  if (impl == FETA_LINEAR_AUTO || impl == FETA_LINEAR_SIMT) {
    const int rc = linear_simt_dx_try(dY, W, dres, mask_src, dX, T, in, out, st);
    if (rc <= 0) return rc;
  }
  if (impl != FETA_LINEAR_MMA) {
    const int rc = linear5_dx_try(dY, W, dres, mask_src, dX, T, in, out, st);
    if (rc <= 0) return rc;
  }
  const size_t wbytes = (size_t)out * (kBN + 8) * sizeof(float);
  if (pick_bm64(T, in)) {
    const size_t smem = (size_t)64 * (out + 4) * sizeof(float) + wbytes;
    FETA_CUDA(cudaFuncSetAttribute(linear_dx_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(T, 64), (unsigned)ceil_div(in, kBN));
    linear_dx_tc_kernel<64><<<grid, 128, smem, st>>>(dY, W, dres, mask_src, dX, T, in, out);
  } else {
    const size_t smem = (size_t)32 * (out + 4) * sizeof(float) + wbytes;
    FETA_CUDA(cudaFuncSetAttribute(linear_dx_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(T, 32), (unsigned)ceil_div(in, kBN));
    linear_dx_tc_kernel<32><<<grid, 64, smem, st>>>(dY, W, dres, mask_src, dX, T, in, out);
  }
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}
