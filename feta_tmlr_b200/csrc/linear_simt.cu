// A6 (layer glue): the encoder layer's projections as fp32 CUDA-core GEMMs built for LATENCY, not throughput.
//
//   Y[T, N]  = act(X[T, K] . W[N, K]^T + b)                       (nn.Linear forward)
//   dX[T, K] = (dY[T, N] . W[N, K]) * [mask > 0] + dres           (its input gradient, ReLU mask / residual fused)
//
// Why: at FeTA's shapes (T = 5k..12k tokens, K, N in {64, 128, 192}) a projection is ~0.1 GFLOP -- every GEMM is
// bound by its own start-up latency, and the step is a dependent chain of ~80 of them.  The library's SIMT sgemm
// walks K in 16-wide stages (global load -> shared -> FMA per stage: ~6.5 us per launch in the step's CUDA graph);
// the tcgen05 kernel (linear_tc5.cu) needs 128-row tiles, i.e. 37 CTAs on 148 SMs at the ZINC shape.  This kernel
// brings the WHOLE reduction dimension of its 32 x 64 output tile into shared memory with one wave of cp.async
// (one memory round trip), runs the products out of shared memory as 128-bit loads + packed FFMA2 (4 x 4 outputs per
// thread, 8 LDS.128 per 32 FFMA2, both operand patterns bank-conflict free), and applies bias / ReLU / ReLU-mask /
// residual-gradient in the epilogue.  Exact fp32 (no TF32 split): results match the library GEMM to rounding.
#include <stdlib.h>

#include "common.cuh"

namespace feta {
namespace lsimt {

constexpr int kBN = 64, kThreads = 128;   // thread (tx = tid % 16, ty = tid / 16): rows ty + 8 i (i < BM / 8), 4 columns

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// MODE 0 (forward): B[kred][n] = W[n0 + n][kred], staged as Ws[n][KR + 4] (reduction index contiguous);
//                   thread columns n = tx + 16 j.
// MODE 1 (dX):      B[kred][n] = W[kred][n0 + n], staged as Ws[kred][kBN + 4] (output index contiguous);
//                   thread columns n = 4 tx + j.
template <int MODE, int kBM>
__global__ void __launch_bounds__(kThreads) linear_simt_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ dres,
                                                              const float* __restrict__ mask_src, float* __restrict__ Y,
                                                              int64_t T, int KR, int NOUT, int ldw, int relu) {
  extern __shared__ __align__(16) float smem[];
  const int lda = KR + 4;                        // padded: rows 4 banks apart -> conflict-free 128-bit row loads
  float* As = smem;                              // [kBM][lda]
  float* Ws = smem + kBM * lda;                  // MODE 0: [kBN][lda];  MODE 1: [KR][kBN + 4]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * kBM;
  const int n0 = blockIdx.y * kBN;
  const int kq = KR >> 2;                        // float4 per A row

  // ---- one wave of asynchronous copies: the whole B panel (a parameter: nothing in the chain writes it, so it is
  // requested while the previous kernel of the chain drains), then -- once that kernel has completed -- the A tile
  pdl_trigger();
  if (MODE == 0) {
    for (int i = tid; i < kBN * kq; i += kThreads) {
      const int n = i / kq, q = i - n * kq;
      cp_async16(Ws + n * lda + 4 * q, W + (int64_t)(n0 + n) * ldw + 4 * q);
    }
  } else {
    constexpr int NQ = kBN / 4;
    for (int i = tid; i < KR * NQ; i += kThreads) {
      const int k = i / NQ, q = i - k * NQ;
      cp_async16(Ws + k * (kBN + 4) + 4 * q, W + (int64_t)k * ldw + n0 + 4 * q);
    }
  }
  pdl_wait();
  for (int i = tid; i < kBM * kq; i += kThreads) {
    const int r = i / kq, q = i - r * kq;
    float* dst = As + r * lda + 4 * q;
    if (row0 + r < T) cp_async16(dst, A + (row0 + r) * KR + 4 * q);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  cp_async_wait_all();
  __syncthreads();

  // Packed fp32x2 FMAs (FFMA2): a 3-register scalar FFMA issues at half rate on this architecture.
  //   forward: a pair = two consecutive reduction indices (both operands are halves of a 128-bit load), i.e. two
  //            interleaved partial sums per output, added at the end;
  //   dX:      a pair = two consecutive output columns (halves of the W row load) times a broadcast A scalar.
  constexpr int RT = kBM / 8;                    // rows per thread
  float acc[RT][4];
  const float* a0 = As + ty * lda;
  if (MODE == 0) {
    float2 p[RT][4];
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) p[i][j] = make_float2(0.f, 0.f);
    const float* w0 = Ws + tx * lda;
#pragma unroll 4
    for (int kk = 0; kk < KR; kk += 4) {
      float4 a[RT], b[4];
#pragma unroll
      for (int i = 0; i < RT; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + (8 * i) * lda + kk);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(w0 + (16 * j) * lda + kk);
#pragma unroll
      for (int i = 0; i < RT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          p[i][j] = __ffma2_rn(make_float2(a[i].x, a[i].y), make_float2(b[j].x, b[j].y), p[i][j]);
          p[i][j] = __ffma2_rn(make_float2(a[i].z, a[i].w), make_float2(b[j].z, b[j].w), p[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = p[i][j].x + p[i][j].y;
  } else {
    float2 p[RT][2];
#pragma unroll
    for (int i = 0; i < RT; ++i) p[i][0] = p[i][1] = make_float2(0.f, 0.f);
    const float* w0 = Ws + 4 * tx;
#pragma unroll 4
    for (int kk = 0; kk < KR; kk += 4) {
      float4 a[RT], b[4];
#pragma unroll
      for (int i = 0; i < RT; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + (8 * i) * lda + kk);
#pragma unroll
      for (int s = 0; s < 4; ++s) b[s] = *reinterpret_cast<const float4*>(w0 + (kk + s) * (kBN + 4));
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        const float as[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
        for (int s2 = 0; s2 < 4; ++s2) {
          const float2 aa = make_float2(as[s2], as[s2]);
          p[i][0] = __ffma2_rn(aa, make_float2(b[s2].x, b[s2].y), p[i][0]);
          p[i][1] = __ffma2_rn(aa, make_float2(b[s2].z, b[s2].w), p[i][1]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < RT; ++i) acc[i][0] = p[i][0].x, acc[i][1] = p[i][0].y, acc[i][2] = p[i][1].x, acc[i][3] = p[i][1].y;
  }

  // ---- epilogue
#pragma unroll
  for (int i = 0; i < RT; ++i) {
    const int64_t row = row0 + ty + 8 * i;
    if (row >= T) continue;
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = n0 + tx + 16 * j;
        float o = acc[i][j];
        if (bias) o += __ldg(bias + col);
        if (relu) o = fmaxf(o, 0.0f);
        Y[row * NOUT + col] = o;
      }
    } else {
      const int col = n0 + 4 * tx;
      float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      if (mask_src) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(mask_src + row * NOUT + col));
        o.x = m.x > 0.f ? o.x : 0.f, o.y = m.y > 0.f ? o.y : 0.f;
        o.z = m.z > 0.f ? o.z : 0.f, o.w = m.w > 0.f ? o.w : 0.f;
      }
      if (dres) {
        const float4 r = __ldg(reinterpret_cast<const float4*>(dres + row * NOUT + col));
        o.x += r.x, o.y += r.y, o.z += r.z, o.w += r.w;
      }
      *reinterpret_cast<float4*>(Y + row * NOUT + col) = o;
    }
  }
}

// Forward projection onto the model width (NOUT = 64 = one column tile, so a CTA owns whole output rows) with the
// layer's degree scale, residual add and LayerNorm in the epilogue:
//   z = res + bscale[row] * (X . W^T + b),   y = (z - mean) * rstd * gamma + beta
// A row's 64 columns sit in the 16 threads of one half warp (4 columns each): mean and variance are two xor-shuffle
// folds.  Replaces the projection launch + the add_layernorm launch of out_proj -> norm1 and linear2 -> norm2.
template <int kBM>
__global__ void __launch_bounds__(kThreads) linear_simt_ln_kernel(
    const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
    const float* __restrict__ res, const float* __restrict__ bscale, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ Y, float* __restrict__ Z, float* __restrict__ mean,
    float* __restrict__ rstd, int64_t T, int KR, float eps) {
  extern __shared__ __align__(16) float smem[];
  const int lda = KR + 4;
  float* As = smem;                              // [kBM][lda]
  float* Ws = smem + kBM * lda;                  // [kBN][lda]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * kBM;
  const int kq = KR >> 2;
  pdl_trigger();
  for (int i = tid; i < kBN * kq; i += kThreads) {
    const int n = i / kq, q = i - n * kq;
    cp_async16(Ws + n * lda + 4 * q, W + (int64_t)n * KR + 4 * q);
  }
  pdl_wait();
  for (int i = tid; i < kBM * kq; i += kThreads) {
    const int r = i / kq, q = i - r * kq;
    float* dst = As + r * lda + 4 * q;
    if (row0 + r < T) cp_async16(dst, A + (row0 + r) * KR + 4 * q);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  cp_async_wait_all();
  __syncthreads();

  constexpr int RT = kBM / 8;
  float2 p[RT][4];
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) p[i][j] = make_float2(0.f, 0.f);
  const float* a0 = As + ty * lda;
  const float* w0 = Ws + tx * lda;
#pragma unroll 4
  for (int kk = 0; kk < KR; kk += 4) {
    float4 a[RT], b[4];
#pragma unroll
    for (int i = 0; i < RT; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + (8 * i) * lda + kk);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(w0 + (16 * j) * lda + kk);
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        p[i][j] = __ffma2_rn(make_float2(a[i].x, a[i].y), make_float2(b[j].x, b[j].y), p[i][j]);
        p[i][j] = __ffma2_rn(make_float2(a[i].z, a[i].w), make_float2(b[j].z, b[j].w), p[i][j]);
      }
  }
  float gm[4], bt[4], bi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int col = tx + 16 * j;
    gm[j] = __ldg(gamma + col), bt[j] = __ldg(beta + col), bi[j] = bias ? __ldg(bias + col) : 0.0f;
  }
#pragma unroll
  for (int i = 0; i < RT; ++i) {
    const int64_t row = row0 + ty + 8 * i;
    const bool live = row < T;
    const float bs = (live && bscale) ? __ldg(bscale + row) : 1.0f;
    float v[4], s = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lin = p[i][j].x + p[i][j].y + bi[j];
      v[j] = live ? fmaf(bs, lin, __ldg(res + row * kBN + tx + 16 * j)) : 0.0f;
      s += v[j];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s * (1.0f / kBN);
    float q = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) q = fmaf(v[j] - mu, v[j] - mu, q);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = rsqrtf(q * (1.0f / kBN) + eps);
    if (!live) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = tx + 16 * j;
      Z[row * kBN + col] = v[j];
      Y[row * kBN + col] = fmaf((v[j] - mu) * rs, gm[j], bt[j]);
    }
    if (tx == 0) mean[row] = mu, rstd[row] = rs;
  }
}

// Backward of the fused launch above, first half: LayerNorm backward over the 64-wide rows as the PROLOGUE of the
// projection's input gradient --
//   dz = rstd (g - mean(g) - xhat mean(g xhat)),  g = dy gamma,  xhat = (z - mean) rstd         (gradient of the residual)
//   dlin = bscale dz                                                            (gradient of the projection's output)
//   dX = (dlin . W) [mask > 0]
// A CTA's A tile holds whole rows, so dz is computed in registers (a row = 16 lanes x 4 columns, two shuffle folds),
// written to shared memory as the GEMM operand and -- by the first column tile only -- to global memory together with
// the per-CTA partial sums of dgamma / dbeta (folded by ln_fold_kernel on the side stream).  Replaces the
// add_layernorm_bwd launch in front of the dX launch of out_proj and linear2.
constexpr int kLnBM = 32;
__global__ void __launch_bounds__(kThreads) lnbwd_dx_kernel(
    const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ bscale,
    const float* __restrict__ W, const float* __restrict__ mask_src, float* __restrict__ dz, float* __restrict__ dlin,
    float* __restrict__ dX, float* __restrict__ partial, int64_t T, int NOUT) {
  extern __shared__ __align__(16) float smem[];
  constexpr int KR = kBN, lda = KR + 4, RT = kLnBM / 8;
  float* As = smem;                              // [kLnBM][lda]  dlin tile
  float* Ws = As + kLnBM * lda;                  // [KR][kBN + 4]
  float* red = Ws + KR * (kBN + 4);              // [2][8][64] column partials of dgamma / dbeta
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = (int64_t)blockIdx.x * kLnBM;
  const int n0 = blockIdx.y * kBN;
  const bool first = blockIdx.y == 0;
  pdl_trigger();
  {
    constexpr int NQ = kBN / 4;
    for (int i = tid; i < KR * NQ; i += kThreads) {
      const int k = i / NQ, q = i - k * NQ;
      cp_async16(Ws + k * (kBN + 4) + 4 * q, W + (int64_t)k * NOUT + n0 + 4 * q);
    }
  }
  const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + tx);
  pdl_wait();
  float dg[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < RT; ++i) {
    const int r = ty + 8 * i;
    const int64_t row = row0 + r;
    const bool live = row < T;
    float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f), z4 = d4;
    float mu = 0.0f, rs = 0.0f, bs = 1.0f;
    if (live) {
      d4 = __ldg(reinterpret_cast<const float4*>(dy + row * KR) + tx);
      z4 = __ldg(reinterpret_cast<const float4*>(z + row * KR) + tx);
      mu = __ldg(mean + row), rs = __ldg(rstd + row);
      if (bscale) bs = __ldg(bscale + row);
    }
    const float d[4] = {d4.x, d4.y, d4.z, d4.w};
    const float xh[4] = {(z4.x - mu) * rs, (z4.y - mu) * rs, (z4.z - mu) * rs, (z4.w - mu) * rs};
    const float g[4] = {d4.x * gm.x, d4.y * gm.y, d4.z * gm.z, d4.w * gm.w};
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s1 += g[j];
      s2 = fmaf(g[j], xh[j], s2);
      dg[j] = fmaf(d[j], xh[j], dg[j]);
      db[j] += d[j];
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 *= (1.0f / KR), s2 *= (1.0f / KR);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = rs * (g[j] - s1 - xh[j] * s2);
    const float4 dl = make_float4(v[0] * bs, v[1] * bs, v[2] * bs, v[3] * bs);
    *reinterpret_cast<float4*>(As + r * lda + 4 * tx) = dl;
    if (first && live) {
      reinterpret_cast<float4*>(dz + row * KR)[tx] = make_float4(v[0], v[1], v[2], v[3]);
      if (dlin) reinterpret_cast<float4*>(dlin + row * KR)[tx] = dl;
    }
  }
  if (first) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[ty * 64 + 4 * tx + j] = dg[j];
      red[512 + ty * 64 + 4 * tx + j] = db[j];
    }
  }
  cp_async_wait_all();
  __syncthreads();
  if (first) {                                   // tid < 64: dgamma column tid; 64..127: dbeta column tid - 64
    const int which = tid >> 6, c = tid & 63;
    float a = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += red[which * 512 + w * 64 + c];
    partial[((size_t)blockIdx.x * 2 + which) * 64 + c] = a;
  }

  float2 p[RT][2];
#pragma unroll
  for (int i = 0; i < RT; ++i) p[i][0] = p[i][1] = make_float2(0.f, 0.f);
  const float* a0 = As + ty * lda;
  const float* w0 = Ws + 4 * tx;
#pragma unroll 4
  for (int kk = 0; kk < KR; kk += 4) {
    float4 a[RT], b[4];
#pragma unroll
    for (int i = 0; i < RT; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + (8 * i) * lda + kk);
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) b[s4] = *reinterpret_cast<const float4*>(w0 + (kk + s4) * (kBN + 4));
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      const float as[4] = {a[i].x, a[i].y, a[i].z, a[i].w};
#pragma unroll
      for (int s2 = 0; s2 < 4; ++s2) {
        const float2 aa = make_float2(as[s2], as[s2]);
        p[i][0] = __ffma2_rn(aa, make_float2(b[s2].x, b[s2].y), p[i][0]);
        p[i][1] = __ffma2_rn(aa, make_float2(b[s2].z, b[s2].w), p[i][1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RT; ++i) {
    const int64_t row = row0 + ty + 8 * i;
    if (row >= T) continue;
    const int col = n0 + 4 * tx;
    float4 o = make_float4(p[i][0].x, p[i][0].y, p[i][1].x, p[i][1].y);
    if (mask_src) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mask_src + row * NOUT + col));
      o.x = m.x > 0.f ? o.x : 0.f, o.y = m.y > 0.f ? o.y : 0.f;
      o.z = m.z > 0.f ? o.z : 0.f, o.w = m.w > 0.f ? o.w : 0.f;
    }
    *reinterpret_cast<float4*>(dX + row * NOUT + col) = o;
  }
}

static bool eligible(int64_t T, int KR, int NOUT, const void* a, const void* w, const void* y, const void* p1,
                     const void* p2) {
  const uintptr_t ptrs = (uintptr_t)a | (uintptr_t)w | (uintptr_t)y | (uintptr_t)p1 | (uintptr_t)p2;
  return T >= 1 && KR >= 4 && KR % 4 == 0 && KR <= 256 && NOUT >= kBN && NOUT % kBN == 0 && NOUT <= 1024 &&
         (ptrs % 16) == 0;
}

template <int MODE, int BM>
static int launch_bm(const float* A, const float* W, const float* bias, const float* dres, const float* mask_src,
                     float* Y, int64_t T, int KR, int NOUT, int ldw, int relu, cudaStream_t st) {
  const size_t smem = ((size_t)BM * (KR + 4) + (MODE == 0 ? (size_t)kBN * (KR + 4) : (size_t)KR * (kBN + 4))) * 4;
  // the opt-in limit only ever grows: lowering it after a launch with a larger tile was captured into a CUDA graph
  // makes a profiler's stand-alone replay of that graph node fail (ncu: LaunchFailed)
  static std::atomic<int> granted{0};
  if ((int)smem > granted.load(std::memory_order_relaxed)) {
    FETA_CUDA(cudaFuncSetAttribute(linear_simt_kernel<MODE, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    granted.store((int)smem, std::memory_order_relaxed);
  }
  dim3 grid((unsigned)ceil_div(T, BM), (unsigned)(NOUT / kBN));
  FETA_CUDA(launch_chain(linear_simt_kernel<MODE, BM>, grid, dim3(kThreads), smem, st, A, W, bias, dres, mask_src, Y, T,
                         KR, NOUT, ldw, relu));
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

template <int MODE>
static int launch(const float* A, const float* W, const float* bias, const float* dres, const float* mask_src, float* Y,
                  int64_t T, int KR, int NOUT, int ldw, int relu, cudaStream_t st) {
  // 64-row tiles halve the re-reads of the W panel (the kernel's L2 -> SM traffic) once 32-row tiles would give every
  // SM more than two CTAs anyway
  const int64_t ctas32 = ceil_div(T, 32) * (NOUT / kBN);
  if (ctas32 > 2 * kNumSMs) return launch_bm<MODE, 64>(A, W, bias, dres, mask_src, Y, T, KR, NOUT, ldw, relu, st);
  return launch_bm<MODE, 32>(A, W, bias, dres, mask_src, Y, T, KR, NOUT, ldw, relu, st);
}

template <int BM>
static int launch_ln_bm(const float* X, const float* W, const float* bias, const float* res, const float* bscale,
                        const float* gamma, const float* beta, float* y, float* z, float* mean, float* rstd, int64_t T,
                        int in, float eps, cudaStream_t st) {
  const size_t smem = ((size_t)BM * (in + 4) + (size_t)kBN * (in + 4)) * 4;
  static std::atomic<int> granted{0};
  if ((int)smem > granted.load(std::memory_order_relaxed)) {
    FETA_CUDA(cudaFuncSetAttribute(linear_simt_ln_kernel<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    granted.store((int)smem, std::memory_order_relaxed);
  }
  FETA_CUDA(launch_chain(linear_simt_ln_kernel<BM>, dim3((unsigned)ceil_div(T, BM)), dim3(kThreads), smem, st, X, W,
                         bias, res, bscale, gamma, beta, y, z, mean, rstd, T, in, eps));
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

}  // namespace lsimt

// y = LayerNorm(res + bscale * (X . W^T + b)); returns 1 when the shape is not eligible (out must be 64)
int linear_simt_ln_try(const float* X, const float* W, const float* bias, const float* res, const float* bscale,
                       const float* gamma, const float* beta, float* y, float* z, float* mean, float* rstd, int64_t T,
                       int in, int out, float eps, cudaStream_t st) {
  if (out != lsimt::kBN || in < 4 || in % 4 != 0 || in > 256 || T < 1) return 1;
  const uintptr_t ptrs = (uintptr_t)X | (uintptr_t)W | (uintptr_t)res | (uintptr_t)y | (uintptr_t)z;
  if (ptrs % 16) return 1;
  // 16-row tiles while 32-row tiles would leave SMs idle (one column tile per row block: T / 32 CTAs)
  if (ceil_div(T, 32) < kNumSMs)
    return lsimt::launch_ln_bm<16>(X, W, bias, res, bscale, gamma, beta, y, z, mean, rstd, T, in, eps, st);
  return lsimt::launch_ln_bm<32>(X, W, bias, res, bscale, gamma, beta, y, z, mean, rstd, T, in, eps, st);
}

// Y[T, out] = act(X[T, in] . W[out, in]^T + b); returns 1 when the shape is not eligible
int linear_simt_fwd_try(const float* X, const float* W, const float* bias, float* Y, int64_t T, int in, int out, int relu,
                        cudaStream_t st) {
  if (!lsimt::eligible(T, in, out, X, W, Y, nullptr, nullptr) || ((uintptr_t)bias % 4)) return 1;
  return lsimt::launch<0>(X, W, bias, nullptr, nullptr, Y, T, in, out, in, relu, st);
}

// dX[T, in] = (dY[T, out] . W[out, in]) * [mask > 0] + dres
int linear_simt_dx_try(const float* dY, const float* W, const float* dres, const float* mask_src, float* dX, int64_t T,
                       int in, int out, cudaStream_t st) {
  if (!lsimt::eligible(T, out, in, dY, W, dX, dres, mask_src)) return 1;
  return lsimt::launch<1>(dY, W, nullptr, dres, mask_src, dX, T, out, in, in, 0, st);
}

}  // namespace feta

extern "C" int feta_lnbwd_linear_dx_blocks(int64_t T) { return (int)feta::ceil_div(T > 0 ? T : 1, feta::lsimt::kLnBM); }

extern "C" int feta_lnbwd_linear_dx(const float* dy, const float* z, const float* mean, const float* rstd,
                                    const float* gamma, const float* bscale, const float* W, const float* mask_src,
                                    float* dz, float* dlin, float* dX, float* partial, int64_t T, int in, int out,
                                    void* stream_) {
  using namespace feta;
  FETA_REQUIRE(T >= 0 && out == lsimt::kBN && in >= 64 && in % 64 == 0 && in <= 1024,
               "lnbwd_linear_dx: unsupported in=%d out=%d (out must be 64, in a multiple of 64)", in, out);
  if (T == 0) return FETA_OK;
  FETA_REQUIRE(dy && z && mean && rstd && gamma && W && dz && dX && partial && ((bscale == nullptr) == (dlin == nullptr)),
               "lnbwd_linear_dx: NULL pointer argument (dlin goes with bscale)");
  const uintptr_t ptrs = (uintptr_t)dy | (uintptr_t)z | (uintptr_t)gamma | (uintptr_t)W | (uintptr_t)mask_src |
                         (uintptr_t)dz | (uintptr_t)dlin | (uintptr_t)dX;
  FETA_REQUIRE((ptrs % 16) == 0, "lnbwd_linear_dx: pointers must be 16-byte aligned");
  const size_t smem = ((size_t)lsimt::kLnBM * (lsimt::kBN + 4) + (size_t)lsimt::kBN * (lsimt::kBN + 4) + 1024) * 4;
  static std::atomic<int> granted{0};
  if ((int)smem > granted.load(std::memory_order_relaxed)) {
    FETA_CUDA(cudaFuncSetAttribute(lsimt::lnbwd_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    granted.store((int)smem, std::memory_order_relaxed);
  }
  dim3 grid((unsigned)ceil_div(T, lsimt::kLnBM), (unsigned)(in / lsimt::kBN));
  FETA_CUDA(launch_chain(lsimt::lnbwd_dx_kernel, grid, dim3(lsimt::kThreads), smem, (cudaStream_t)stream_, dy, z, mean,
                         rstd, gamma, bscale, W, mask_src, dz, dlin, dX, partial, T, in));
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_linear_layernorm_simt_supported(int in, int out) {
  return out == 64 && in >= 64 && in % 64 == 0 && in <= 256;
}

extern "C" int feta_linear_simt_supported(int in, int out) {
  return in >= 64 && out >= 64 && in % 64 == 0 && out % 64 == 0 && in <= 256 && out <= 256;
}
