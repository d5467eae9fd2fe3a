// A6 (layer glue): the two token-axis reductions of DiffTransformerEncoderLayer's backward that
// the stock libraries under-parallelise at FeTA's shapes (T = Nmax*B ~ 5k tokens, d = 64..192):
//   * weight gradients  dW[out, in] = sum_t dY[t, out] X[t, in]  (+ db = sum_t dY[t])  -- cuBLAS picks a
//     54-CTA SIMT sgemm taking 39 us (profiles/r1a_launches_eager_zinc.md); here the token axis is
//     split over ~T/128 CTAs per output tile, partials reduced deterministically in a second pass;
//   * residual-add + LayerNorm forward/backward with the gamma/beta gradients reduced across all
//     SMs (PyTorch's GammaBetaBackward runs on 2 CTAs, 39 us).
// These replace `F.linear` backward / `nn.LayerNorm` inside the layer the reference imports at
// transformer/models.py:4 (residual + norm1 / FFN + norm2 of the GraphiT layer).
#include "common.cuh"

namespace feta {

constexpr int kWgTile = 64, kWgTT = 32, kWgThreads = 256;
constexpr int kWgPitch = kWgTile + 8;   // smem row pitch: fragment loads (4 token rows x 8 columns per warp) hit 32 banks

// 3xTF32 (hi.hi + hi.lo + lo.hi, fp32-grade): the tensor core reads the top 19 bits of an operand, so "hi" is the
// value itself and the remainder is exact in fp32 (same scheme as csrc/cheb_lane.cu)
__device__ __forceinline__ void wg_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x);
  lo = __float_as_uint(x - __uint_as_float(hi & 0xFFFFE000u));
}
__device__ __forceinline__ void wg_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One CTA = one 64 x 64 tile of dW over one slice of `t_per_cta` tokens.  The contraction runs over tokens on the
// tensor cores: M = output channel (A[m][k] = dY[t][o], read transposed out of the token-major staging tile),
// N = input channel (B[k][n] = X[t][i]), K = 8 tokens per step; warp w owns output rows 16 (w & 3) .. + 15 and input
// columns 32 (w >> 2) .. + 31.  The round-2 launch list had the CUDA-core version of this kernel (16 FFMA per two
// LDS.128, 3840 warp instructions per slice) as the largest single share of the step's GPU time, competing for issue
// slots with the layer chain it runs beside; this one issues ~1200.  db rides along as one more accumulator tile
// against an all-ones B fragment.
__global__ void __launch_bounds__(kWgThreads) wgrad_partial_kernel(const float* __restrict__ dY,
                                                                  const float* __restrict__ X,
                                                                  float* __restrict__ partial,
                                                                  float* __restrict__ partial_db, int T, int out,
                                                                  int in, int t_per_cta, int in_tiles) {
  __shared__ __align__(16) float sA[kWgTT][kWgPitch];
  __shared__ __align__(16) float sB[kWgTT][kWgPitch];
  const int tile = blockIdx.x, s = blockIdx.y;
  const int o0 = (tile / in_tiles) * kWgTile, i0 = (tile % in_tiles) * kWgTile;
  const int t0 = s * t_per_cta, t1 = min(T, t0 + t_per_cta);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tq = lane & 3;
  const int mo = (warp & 3) * 16, nb = (warp >> 2) * 32;
  const bool want_db = partial_db != nullptr && i0 == 0 && nb == 0;   // warp-uniform
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0f;
  float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
  const uint32_t one = __float_as_uint(1.0f);
  // register-staged prefetch: the next 32-token tile is in flight while the current one is consumed
  float4 ra[2], rb[2];
  auto fetch = [&](int tt0) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int idx = threadIdx.x + kWgThreads * j;
      const int row = idx >> 4, c4 = (idx & 15) * 4;
      const int t = tt0 + row;
      ra[j] = rb[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < t1) {
        if (o0 + c4 < out) ra[j] = __ldg(reinterpret_cast<const float4*>(dY + (size_t)t * out + o0 + c4));
        if (i0 + c4 < in) rb[j] = __ldg(reinterpret_cast<const float4*>(X + (size_t)t * in + i0 + c4));
      }
    }
  };
  fetch(t0);
  for (int tt0 = t0; tt0 < t1; tt0 += kWgTT) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int idx = threadIdx.x + kWgThreads * j;
      const int row = idx >> 4, c4 = (idx & 15) * 4;
      *reinterpret_cast<float4*>(&sA[row][c4]) = ra[j];
      *reinterpret_cast<float4*>(&sB[row][c4]) = rb[j];
    }
    __syncthreads();
    if (tt0 + kWgTT < t1) fetch(tt0 + kWgTT);
#pragma unroll
    for (int ks = 0; ks < kWgTT / 8; ++ks) {
      const float* pa = &sA[8 * ks + tq][mo + g];      // A[m][k]: rows g, g + 8 <-> channels, k = tq, tq + 4 <-> tokens
      uint32_t ah[4], al[4];
      wg_split(pa[0], ah[0], al[0]);
      wg_split(pa[8], ah[1], al[1]);
      wg_split(pa[4 * kWgPitch], ah[2], al[2]);
      wg_split(pa[4 * kWgPitch + 8], ah[3], al[3]);
      const float* pb = &sB[8 * ks + tq][nb + g];      // B[k][n]: k = tq, tq + 4 <-> tokens, n = g <-> channel
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        uint32_t bh0, bl0, bh1, bl1;
        wg_split(pb[8 * nt], bh0, bl0);
        wg_split(pb[4 * kWgPitch + 8 * nt], bh1, bl1);
        wg_mma(acc[nt], ah, bh0, bh1);
        wg_mma(acc[nt], ah, bl0, bl1);
        wg_mma(acc[nt], al, bh0, bh1);
      }
      if (want_db) {
        wg_mma(dbacc, ah, one, one);
        wg_mma(dbacc, al, one, one);
      }
    }
    __syncthreads();
  }
  // C fragment: (row g, cols 2 tq, 2 tq + 1), (row g + 8, same cols)
  float* pt = partial + (size_t)s * out * in;
  const int oa = o0 + mo + g, ob = oa + 8;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int i = i0 + nb + 8 * nt + 2 * tq;
    if (i < in) {
      if (oa < out) *reinterpret_cast<float2*>(pt + (size_t)oa * in + i) = make_float2(acc[nt][0], acc[nt][1]);
      if (ob < out) *reinterpret_cast<float2*>(pt + (size_t)ob * in + i) = make_float2(acc[nt][2], acc[nt][3]);
    }
  }
  if (want_db && tq == 0) {     // every column of the ones-product holds sum_t dY[t][o]
    if (oa < out) partial_db[(size_t)s * out + oa] = dbacc[0];
    if (ob < out) partial_db[(size_t)s * out + ob] = dbacc[2];
  }
}

// dW[e] = sum_s partial[s, e] (and db[e] = sum_s pdb[s, e], same launch) over float4 elements; slices summed in
// order (deterministic), loads unrolled
__global__ void __launch_bounds__(128) slices_reduce4_kernel(const float4* __restrict__ partial, int S, int64_t n4,
                                                            float4* __restrict__ out,
                                                            const float4* __restrict__ partial2, int64_t n4b,
                                                            float4* __restrict__ out2) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n4) {   // tail threads: the bias slices
    e -= n4;
    if (e >= n4b) return;
    partial = partial2, out = out2, n4 = n4b;
  }
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  int k = 0;
  for (; k + 8 <= S; k += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldg(partial + (size_t)(k + u) * n4 + e);
#pragma unroll
    for (int u = 0; u < 8; ++u) a.x += v[u].x, a.y += v[u].y, a.z += v[u].z, a.w += v[u].w;
  }
  for (; k < S; ++k) {
    const float4 v = __ldg(partial + (size_t)k * n4 + e);
    a.x += v.x, a.y += v.y, a.z += v.z, a.w += v.w;
  }
  out[e] = a;
}

// ---------------------------------------------------------------- residual add + LayerNorm ----
constexpr int kLnWarps = 8, kLnMaxPerLane = 8;  // D <= 256

__global__ void __launch_bounds__(kLnWarps * 32) add_layernorm_fwd_kernel(
    const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ bscale,
    const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ z,
    float* __restrict__ mean, float* __restrict__ rstd, int64_t T, int D, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  pdl_trigger();
  pdl_wait();          // a / b come from the previous kernel of the chain
  if (row >= T) return;
  float v[kLnMaxPerLane];
  float s = 0.0f;
  const float bs = bscale ? bscale[row] : 1.0f;
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j) {
    const int c = lane + 32 * j;
    v[j] = 0.0f;
    if (c < D) {
      v[j] = a[row * D + c] + (b ? bs * b[row * D + c] : 0.0f);
      s += v[j];
    }
  }
  const float mu = warp_sum(s) / (float)D;
  float q = 0.0f;
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j) {
    const int c = lane + 32 * j;
    if (c < D) q = fmaf(v[j] - mu, v[j] - mu, q);
  }
  const float rs = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j) {
    const int c = lane + 32 * j;
    if (c < D) {
      z[row * D + c] = v[j];
      y[row * D + c] = (v[j] - mu) * rs * gamma[c] + beta[c];
    }
  }
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// dz (critical path) + per-CTA partial sums of dgamma / dbeta.  PER = ceil(D / 32) columns per lane.  With
// FOLD the last CTA to finish also folds the partials (fixed order, self-re-arming counter); without it the
// fold is a separate launch (ln_fold_kernel) that the host puts on the weight-gradient side stream, so the
// parameter gradients never sit on the critical path of the backward pass.
template <int PER, bool FOLD>
__global__ void __launch_bounds__(kLnWarps * 32) add_layernorm_bwd_kernel(
    const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ bscale,
    float* __restrict__ dz, float* __restrict__ db_scaled, float* __restrict__ partial /* [grid, 2, D] */,
    float* __restrict__ dgamma, float* __restrict__ dbeta, int32_t* __restrict__ counter, int64_t T, int D) {
  __shared__ float sg[kLnWarps][PER * 32];
  __shared__ float sb[kLnWarps][PER * 32];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float dg[PER], db[PER], gm[PER];
  pdl_trigger();
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    dg[j] = db[j] = 0.0f;
    const int c = lane + 32 * j;
    gm[j] = c < D ? gamma[c] : 0.0f;             // a parameter: not written by the chain
  }
  pdl_wait();          // dy comes from the previous kernel of the chain
  for (int64_t row = (int64_t)blockIdx.x * kLnWarps + warp; row < T; row += (int64_t)gridDim.x * kLnWarps) {
    const float mu = mean[row], rs = rstd[row];
    float g[PER], xh[PER];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int c = lane + 32 * j;
      g[j] = xh[j] = 0.0f;
      if (c < D) {
        const float d = dy[row * D + c];
        xh[j] = (z[row * D + c] - mu) * rs;
        g[j] = d * gm[j];
        s1 += g[j];
        s2 = fmaf(g[j], xh[j], s2);
        dg[j] = fmaf(d, xh[j], dg[j]);
        db[j] += d;
      }
    }
    // the two row sums travel through the shuffles together
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 /= (float)D;
    s2 /= (float)D;
    const float bs = db_scaled ? bscale[row] : 0.0f;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int c = lane + 32 * j;
      if (c < D) {
        const float v = rs * (g[j] - s1 - xh[j] * s2);
        dz[row * D + c] = v;
        if (db_scaled) db_scaled[row * D + c] = v * bs;   // gradient of the scaled branch
      }
    }
  }
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    sg[warp][lane + 32 * j] = dg[j];
    sb[warp][lane + 32 * j] = db[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float a = 0.0f, b = 0.0f;
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) {
      a += sg[w][c];
      b += sb[w][c];
    }
    partial[((size_t)blockIdx.x * 2 + 0) * D + c] = a;
    partial[((size_t)blockIdx.x * 2 + 1) * D + c] = b;
  }
  if constexpr (!FOLD) return;
  // last block folds the per-block partials (fixed order) and re-arms the counter
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int done = atomicAdd(counter, 1);
    s_last = (done == (int)gridDim.x - 1);
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // one warp per (column, which) item, lanes stride the CTA partials (same tree as ln_fold_kernel)
  for (int item = warp; item < 2 * D; item += kLnWarps) {
    const int which = item / D, c = item - which * D;
    float a = 0.0f;
    for (int k = lane; k < (int)gridDim.x; k += 32) a += __ldcg(partial + ((size_t)k * 2 + which) * D + c);
    a = warp_sum(a);
    if (lane == 0) (which ? dbeta : dgamma)[c] = a;
  }
}

// dgamma[c] = sum_k partial[k, 0, c], dbeta[c] = sum_k partial[k, 1, c]: one warp per (column, which), lanes
// stride the CTA partials, fixed-shape tree => deterministic
__global__ void __launch_bounds__(256) ln_fold_kernel(const float* __restrict__ partial, int nblk, int D,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);   // 0 .. 2D-1
  if (item >= 2 * D) return;
  const int which = item / D, c = item - which * D;
  float a = 0.0f;
  for (int k = lane; k < nblk; k += 32) a += __ldg(partial + ((size_t)k * 2 + which) * D + c);
  a = warp_sum(a);
  if (lane == 0) (which ? dbeta : dgamma)[c] = a;
}

}  // namespace feta

using namespace feta;

// tokens per CTA (= per slice of the token axis).  Measured in the step (round 2, tensor-core kernel): ZINC shape
// (T = 4.7k) 128 tokens 101.9k graphs/s, 192 99.3k, 256 94.2k; PATTERN shape (T = 12k) 192 54.1k, 128 53.0k -- short
// slices put more CTAs on a small batch, long ones keep the second pass short on a large one
static inline int wg_tokens(int64_t T) { return T <= 8192 ? 128 : 192; }
extern "C" int feta_linear_wgrad_slices(int64_t T) { return (int)ceil_div(T > 0 ? T : 1, wg_tokens(T)); }

extern "C" int feta_linear_wgrad(const float* dY, const float* X, float* dW, float* db, float* partial,
                                 size_t partial_floats, int32_t* counters, int64_t T, int out, int in,
                                 void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(T >= 0 && out >= 4 && in >= 4 && out % 4 == 0 && in % 4 == 0,
               "linear_wgrad: needs out, in multiples of 4 (got %d, %d)", out, in);
  (void)counters;
  FETA_REQUIRE(dW && partial && (T == 0 || (dY && X)), "linear_wgrad: NULL pointer argument");
  FETA_REQUIRE(((uintptr_t)dW % 16 == 0) && (!db || (uintptr_t)db % 16 == 0), "linear_wgrad: dW/db must be 16-byte aligned");
  FETA_REQUIRE(((uintptr_t)dY % 16 == 0) && ((uintptr_t)X % 16 == 0) && ((uintptr_t)partial % 16 == 0),
               "linear_wgrad: pointers must be 16-byte aligned");
  const int S = feta_linear_wgrad_slices(T);
  if (partial_floats < (size_t)S * ((size_t)out * in + out)) {
    set_last_error("linear_wgrad: partial buffer too small (%zu < %zu floats)", partial_floats,
                   (size_t)S * ((size_t)out * in + out));
    return FETA_EWORKSPACE;
  }
  float* pdb = db ? partial + (size_t)S * out * in : nullptr;
  const int out_tiles = (int)ceil_div(out, kWgTile), in_tiles = (int)ceil_div(in, kWgTile);
  dim3 grid((unsigned)(out_tiles * in_tiles), (unsigned)S);
  wgrad_partial_kernel<<<grid, kWgThreads, 0, st>>>(dY, X, partial, pdb, (int)T, out, in, wg_tokens(T), in_tiles);
  FETA_LAUNCH_CHECK();
  const int64_t n4 = (int64_t)out * in / 4, n4b = db ? out / 4 : 0;
  slices_reduce4_kernel<<<(unsigned)ceil_div(n4 + n4b, 128), 128, 0, st>>>(
      reinterpret_cast<const float4*>(partial), S, n4, reinterpret_cast<float4*>(dW),
      reinterpret_cast<const float4*>(pdb), n4b, reinterpret_cast<float4*>(db));
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_add_layernorm_fwd(const float* a, const float* b, const float* bscale, const float* gamma,
                                      const float* beta, float* y, float* z, float* mean, float* rstd, int64_t T, int D,
                                      float eps, void* stream_) {
  FETA_REQUIRE(T >= 0 && D >= 1 && D <= kLnMaxPerLane * 32, "add_layernorm: D=%d not in [1, %d]", D,
               kLnMaxPerLane * 32);
  if (T == 0) return FETA_OK;
  FETA_REQUIRE(a && gamma && beta && y && z && mean && rstd && (b || !bscale),
               "add_layernorm_fwd: NULL pointer argument");
  FETA_CUDA(launch_chain(add_layernorm_fwd_kernel, dim3((unsigned)ceil_div(T, kLnWarps)), dim3(kLnWarps * 32), 0,
                         (cudaStream_t)stream_, a, b, bscale, gamma, beta, y, z, mean, rstd, T, D, eps));
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_add_layernorm_bwd_blocks(int64_t T) {
  int64_t b = ceil_div(T > 0 ? T : 1, kLnWarps * 2);   // two rows per warp: enough CTAs to fill the GPU
  return (int)(b < 4 * kNumSMs ? b : 4 * kNumSMs);
}

template <int PER>
static void launch_ln_bwd(bool fold, int nblk, cudaStream_t st, const float* dy, const float* z, const float* mean,
                          const float* rstd, const float* gamma, const float* bscale, float* dz, float* db_scaled,
                          float* partial, float* dgamma, float* dbeta, int32_t* counter, int64_t T, int D) {
  if (fold)
    launch_chain(add_layernorm_bwd_kernel<PER, true>, dim3(nblk), dim3(kLnWarps * 32), 0, st, dy, z, mean, rstd, gamma,
                 bscale, dz, db_scaled, partial, dgamma, dbeta, counter, T, D);
  else
    launch_chain(add_layernorm_bwd_kernel<PER, false>, dim3(nblk), dim3(kLnWarps * 32), 0, st, dy, z, mean, rstd, gamma,
                 bscale, dz, db_scaled, partial, dgamma, dbeta, counter, T, D);
}

extern "C" int feta_add_layernorm_bwd(const float* dy, const float* z, const float* mean, const float* rstd,
                                      const float* gamma, const float* bscale, float* dz, float* db_scaled,
                                      float* dgamma, float* dbeta, float* partial, int32_t* counter, int64_t T, int D,
                                      void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(T >= 0 && D >= 1 && D <= kLnMaxPerLane * 32, "add_layernorm: D=%d not in [1, %d]", D,
               kLnMaxPerLane * 32);
  const bool fold = dgamma != nullptr;
  FETA_REQUIRE(partial && (!fold || (dbeta && counter)) && (T == 0 || (dy && z && mean && rstd && gamma && dz)),
               "add_layernorm_bwd: NULL pointer argument");
  FETA_REQUIRE(!db_scaled || bscale, "add_layernorm_bwd: db_scaled needs bscale");
  const int nblk = feta_add_layernorm_bwd_blocks(T);
  if (D <= 32) launch_ln_bwd<1>(fold, nblk, st, dy, z, mean, rstd, gamma, bscale, dz, db_scaled, partial, dgamma, dbeta, counter, T, D);
  else if (D <= 64) launch_ln_bwd<2>(fold, nblk, st, dy, z, mean, rstd, gamma, bscale, dz, db_scaled, partial, dgamma, dbeta, counter, T, D);
  else if (D <= 128) launch_ln_bwd<4>(fold, nblk, st, dy, z, mean, rstd, gamma, bscale, dz, db_scaled, partial, dgamma, dbeta, counter, T, D);
  else launch_ln_bwd<8>(fold, nblk, st, dy, z, mean, rstd, gamma, bscale, dz, db_scaled, partial, dgamma, dbeta, counter, T, D);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_add_layernorm_bwd_fold(const float* partial, int64_t T, int D, float* dgamma, float* dbeta,
                                           void* stream_) {
  FETA_REQUIRE(partial && dgamma && dbeta && D >= 1 && D <= kLnMaxPerLane * 32, "add_layernorm_bwd_fold: bad argument");
  const int nblk = feta_add_layernorm_bwd_blocks(T);
  ln_fold_kernel<<<(unsigned)ceil_div(2 * D, 8), 256, 0, (cudaStream_t)stream_>>>(partial, nblk, D, dgamma, dbeta);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

// the same fold over an explicit number of per-CTA partials (feta_lnbwd_linear_dx writes one per 32-row tile)
extern "C" int feta_ln_fold(const float* partial, int nblk, int D, float* dgamma, float* dbeta, void* stream_) {
  FETA_REQUIRE(partial && dgamma && dbeta && nblk >= 0 && D >= 1 && D <= kLnMaxPerLane * 32, "ln_fold: bad argument");
  ln_fold_kernel<<<(unsigned)ceil_div(2 * D, 8), 256, 0, (cudaStream_t)stream_>>>(partial, nblk, D, dgamma, dbeta);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

