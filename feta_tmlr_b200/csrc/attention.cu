// A6: kernel-biased attention core, fp32-exact CUDA-core path (see include/feta_b200.h).
//
// Replaces the un-fused bmm -> masked_fill -> max -> exp -> *pe -> /clamp(sum) -> bmm chain of the
// (missing) DiffTransformerEncoderLayer that transformer/models.py:4 imports; contract from
// models.py:166-167.  One launch computes S, the kernel-biased normalisation, writes the
// attention matrix the caller needs (models.py:173 consumes it) and O = P V per head, with K/V
// of the (graph, head) staged once in shared memory.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace feta {

// csrc/attention_tc.cu: tcgen05 / TMEM forward (3xTF32); 1 = shape not eligible
int attn_fwd_tc_try(const float* q, const float* k, const float* v, int64_t sn, int64_t sb, const float* pe,
                    const uint8_t* mask, float* attn, float* o_heads, int64_t osn, int64_t osb, float* rowflag, int B,
                    int H, int nmax, int dh, float scale, cudaStream_t st);

constexpr int kAttnThreads = 256;
constexpr int kAttnWarps = kAttnThreads / 32;
constexpr int kRowsPerCta = 32;  // forward: query rows per CTA

// n_eff = 1 + index of the last un-masked position of graph b (0 if all masked)
__device__ __forceinline__ int block_n_eff(const uint8_t* __restrict__ mk, int nmax, int* s_neff) {
  if (threadIdx.x == 0) *s_neff = 0;
  __syncthreads();
  int loc = 0;
  for (int j = threadIdx.x; j < nmax; j += blockDim.x)
    if (mk[j] == 0) loc = j + 1;
  if (loc > 0) atomicMax(s_neff, loc);
  __syncthreads();
  return *s_neff;
}

template <int DH>
__device__ __forceinline__ float slice_reduce(float v) {
  // lanes are (c = lane % DH, slice = lane / DH); sum over slices
#pragma unroll
  for (int o = 16; o >= DH; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// first key position whose score equals the row maximum (lane j + 32 ch holds sraw[ch])
template <int NCH>
__device__ __forceinline__ int clamped_argmax(const float (&sraw)[NCH], float m, int lane) {
  int best = 0x7fffffff;
#pragma unroll
  for (int ch = 0; ch < NCH; ++ch)
    if (sraw[ch] == m && m != -INFINITY) best = min(best, lane + 32 * ch);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  return best == 0x7fffffff ? 0 : best;
}

// ------------------------------------------------------------------ forward -------------
template <int DH, int NCH>
__global__ void __launch_bounds__(kAttnThreads) attn_fwd_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t sn, int64_t sb,
    const float* __restrict__ pe, const uint8_t* __restrict__ mask, float* __restrict__ attn,
    float* __restrict__ o_heads, int64_t osn, int64_t osb, float* __restrict__ rowflag, int H, int nmax,
    float scale, int rows_per_cta, const float* __restrict__ drop, float* __restrict__ attn_post) {
  extern __shared__ float smem[];
  __shared__ int s_neff;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int i0 = blockIdx.x * rows_per_cta;
  const uint8_t* mk = mask + (size_t)b * nmax;
  const int n = block_n_eff(mk, nmax, &s_neff);
  const int npad = nmax | 1;
  float* Kt = smem;                         // [DH][npad]
  float* Vs = Kt + (size_t)DH * npad;       // [nmax][DH]
  float* prow = Vs + (size_t)nmax * DH;     // [warps][npad]
  float* pen = prow + (size_t)kAttnWarps * npad;  // [nmax] 0 or -inf key penalty

  const float* kb = k + (int64_t)b * sb + h * DH;
  const float* vb = v + (int64_t)b * sb + h * DH;
  for (int idx = threadIdx.x; idx < n * DH; idx += blockDim.x) {
    const int j = idx / DH, c = idx - j * DH;
    Kt[c * npad + j] = __ldg(kb + (int64_t)j * sn + c);
    Vs[idx] = __ldg(vb + (int64_t)j * sn + c);
  }
  for (int j = threadIdx.x; j < nmax; j += blockDim.x) pen[j] = mk[j] ? -INFINITY : 0.0f;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* pr = prow + (size_t)warp * npad;
  for (int ii = warp; ii < rows_per_cta; ii += kAttnWarps) {
    const int i = i0 + ii;
    if (i >= nmax) break;
    float* arow = attn + (((size_t)b * H + h) * nmax + i) * nmax;
    float* orow = o_heads + (int64_t)i * osn + (int64_t)b * osb + h * DH;
    if (mk[i]) {  // padded query: defined as zero (never consumed by the model)
      for (int j = lane; j < nmax; j += 32) arow[j] = 0.0f;
      if (attn_post)
        for (int j = lane; j < nmax; j += 32) attn_post[(((size_t)b * H + h) * nmax + i) * nmax + j] = 0.0f;
      if (lane < DH) orow[lane] = 0.0f;
      if (DH > 32 && lane + 32 < DH) orow[lane + 32] = 0.0f;
      if (lane == 0) rowflag[((size_t)b * H + h) * nmax + i] = 0.0f;
      continue;
    }
    float qr[DH];
    const float* qp = q + (int64_t)i * sn + (int64_t)b * sb + h * DH;
#pragma unroll
    for (int c = 0; c < DH; ++c) qr[c] = __ldg(qp + c) * scale;  // q * scaling before the product
    float s[NCH], sraw[NCH];
    float m = -INFINITY;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int j = lane + 32 * ch;
      float a = -INFINITY;
      if (j < n) {
        a = 0.0f;
#pragma unroll
        for (int c = 0; c < DH; ++c) a = fmaf(qr[c], Kt[c * npad + j], a);
        a += pen[j];
      }
      s[ch] = a;
      sraw[ch] = a;
      m = fmaxf(m, a);
    }
    m = warp_max(m);
    float sum = 0.0f;
    const float* perow = pe ? pe + ((size_t)b * nmax + i) * nmax : nullptr;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int j = lane + 32 * ch;
      float e = 0.0f;
      if (j < n && s[ch] != -INFINITY) {
        e = expf(s[ch] - m);
        if (perow) e *= __ldg(perow + j);
      }
      s[ch] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float denom = fmaxf(sum, 1e-6f);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int j = lane + 32 * ch;
      const float p = s[ch] / denom;
      float pd = p;        // attention-weight dropout: P V uses P * drop (0 or 1/(1-p)); `attn` keeps P for backward
      if (drop != nullptr && j < nmax) {
        const size_t at = (((size_t)b * H + h) * nmax + i) * nmax + j;
        pd = p * __ldg(drop + at);
        attn_post[at] = pd;
      }
      if (j < nmax) arow[j] = p;
      if (j < n) pr[j] = pd;
    }
    {
      // rowflag: 1 = normalised row; -(1 + argmax_j S_ij) = row under the clamp (constant denominator, but the row
      // maximum still carries a gradient: oracle/layers.py:61-66 does not detach it); 0 = padded query
      float rf = 1.0f;
      if (!(sum > 1e-6f)) rf = -(float)(1 + clamped_argmax<NCH>(sraw, m, lane));
      if (lane == 0) rowflag[((size_t)b * H + h) * nmax + i] = rf;
    }
    __syncwarp();
    if (DH <= 32) {
      constexpr int NS = DH <= 32 ? 32 / DH : 1;
      const int c = lane % DH, js = lane / DH;
      float o = 0.0f;
      for (int j = js; j < n; j += NS) o = fmaf(pr[j], Vs[j * DH + c], o);
      o = slice_reduce<(DH <= 32 ? DH : 32)>(o);
      if (js == 0) orow[c] = o;
    } else {
      float o0 = 0.0f, o1 = 0.0f;
      for (int j = 0; j < n; ++j) {
        const float p = pr[j];
        o0 = fmaf(p, Vs[j * DH + lane], o0);
        o1 = fmaf(p, Vs[j * DH + lane + 32], o1);
      }
      orow[lane] = o0;
      orow[lane + 32] = o1;
    }
    __syncwarp();
  }
}

// Forward for graphs of > 64 nodes at dh 8 / 16 (PATTERN / CLUSTER / molhiv shapes): K and V row-major with a
// padded stride (one 128-bit load per 4 FMAs in QK^T and PV instead of one 32-bit load per FMA) -- the kernel
// is instruction bound, so this is where its time goes.  Same contract and row mapping as attn_fwd_kernel.
template <int DH, int NCH>
__global__ void __launch_bounds__(kAttnThreads) attn_fwd_tiled_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t sn, int64_t sb,
    const float* __restrict__ pe, const uint8_t* __restrict__ mask, float* __restrict__ attn,
    float* __restrict__ o_heads, int64_t osn, int64_t osb, float* __restrict__ rowflag, int H, int nmax,
    float scale, int rows_per_cta, const float* __restrict__ drop, float* __restrict__ attn_post) {
  extern __shared__ float smem[];
  __shared__ int s_neff;
  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int i0 = blockIdx.x * rows_per_cta;
  const uint8_t* mk = mask + (size_t)b * nmax;
  const int n = block_n_eff(mk, nmax, &s_neff);
  const int npad = nmax | 1;
  constexpr int LD = DH + 4, C4 = DH / 4;   // rows padded to DH + 4 floats: conflict-free 128-bit row loads
  float* Ks = smem;                         // [nmax][LD]
  float* Vs = Ks + (size_t)nmax * LD;       // [nmax][LD]
  float* prow = Vs + (size_t)nmax * LD;     // [warps][npad]
  float* pen = prow + (size_t)kAttnWarps * npad;  // [nmax] 0 or -inf key penalty

  const float* kb = k + (int64_t)b * sb + h * DH;
  const float* vb = v + (int64_t)b * sb + h * DH;
  for (int idx = threadIdx.x; idx < n * C4; idx += blockDim.x) {
    const int j = idx / C4, c = (idx - j * C4) * 4;
    *reinterpret_cast<float4*>(Ks + j * LD + c) = __ldg(reinterpret_cast<const float4*>(kb + (int64_t)j * sn + c));
    *reinterpret_cast<float4*>(Vs + j * LD + c) = __ldg(reinterpret_cast<const float4*>(vb + (int64_t)j * sn + c));
  }
  for (int j = threadIdx.x; j < nmax; j += blockDim.x) pen[j] = mk[j] ? -INFINITY : 0.0f;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* pr = prow + (size_t)warp * npad;
  for (int ii = warp; ii < rows_per_cta; ii += kAttnWarps) {
    const int i = i0 + ii;
    if (i >= nmax) break;
    float* arow = attn + (((size_t)b * H + h) * nmax + i) * nmax;
    float* orow = o_heads + (int64_t)i * osn + (int64_t)b * osb + h * DH;
    if (mk[i]) {  // padded query: defined as zero (never consumed by the model)
      for (int j = lane; j < nmax; j += 32) arow[j] = 0.0f;
      if (attn_post)
        for (int j = lane; j < nmax; j += 32) attn_post[(((size_t)b * H + h) * nmax + i) * nmax + j] = 0.0f;
      if (lane < DH) orow[lane] = 0.0f;
      if (lane == 0) rowflag[((size_t)b * H + h) * nmax + i] = 0.0f;
      continue;
    }
    float qr[DH];
    const float* qp = q + (int64_t)i * sn + (int64_t)b * sb + h * DH;
#pragma unroll
    for (int c = 0; c < DH; ++c) qr[c] = __ldg(qp + c) * scale;  // q * scaling before the product
    float s[NCH], sraw[NCH];
    float m = -INFINITY;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int j = lane + 32 * ch;
      float a = -INFINITY;
      if (j < n) {
        a = 0.0f;
#pragma unroll
        for (int c = 0; c < C4; ++c) {
          const float4 t4 = *reinterpret_cast<const float4*>(Ks + j * LD + 4 * c);
          a = fmaf(qr[4 * c], t4.x, a), a = fmaf(qr[4 * c + 1], t4.y, a);
          a = fmaf(qr[4 * c + 2], t4.z, a), a = fmaf(qr[4 * c + 3], t4.w, a);
        }
        a += pen[j];
      }
      s[ch] = a;
      sraw[ch] = a;
      m = fmaxf(m, a);
    }
    m = warp_max(m);
    float sum = 0.0f;
    const float* perow = pe ? pe + ((size_t)b * nmax + i) * nmax : nullptr;
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int j = lane + 32 * ch;
      float e = 0.0f;
      if (j < n && s[ch] != -INFINITY) {
        e = expf(s[ch] - m);
        if (perow) e *= __ldg(perow + j);
      }
      s[ch] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float denom = fmaxf(sum, 1e-6f);
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const int j = lane + 32 * ch;
      const float p = s[ch] / denom;
      float pd = p;        // attention-weight dropout: P V uses P * drop (0 or 1/(1-p)); `attn` keeps P for backward
      if (drop != nullptr && j < nmax) {
        const size_t at = (((size_t)b * H + h) * nmax + i) * nmax + j;
        pd = p * __ldg(drop + at);
        attn_post[at] = pd;
      }
      if (j < nmax) arow[j] = p;
      if (j < n) pr[j] = pd;
    }
    {
      // rowflag: 1 = normalised row; -(1 + argmax_j S_ij) = row under the clamp (constant denominator, but the row
      // maximum still carries a gradient: oracle/layers.py:61-66 does not detach it); 0 = padded query
      float rf = 1.0f;
      if (!(sum > 1e-6f)) rf = -(float)(1 + clamped_argmax<NCH>(sraw, m, lane));
      if (lane == 0) rowflag[((size_t)b * H + h) * nmax + i] = rf;
    }
    __syncwarp();
    {  // O_i = sum_j P[j] V[j]: lanes = (4-channel group, key slice), one 128-bit V load per 4 FMAs
      constexpr int NS = 32 / C4;
      const int lc = lane % C4, js = lane / C4;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = js; j < n; j += NS) {
        const float w = pr[j];
        const float4 t4 = *reinterpret_cast<const float4*>(Vs + j * LD + 4 * lc);
        o.x = fmaf(w, t4.x, o.x), o.y = fmaf(w, t4.y, o.y), o.z = fmaf(w, t4.z, o.z), o.w = fmaf(w, t4.w, o.w);
      }
#pragma unroll
      for (int sh = 16; sh >= C4; sh >>= 1) {
        o.x += __shfl_xor_sync(0xffffffffu, o.x, sh), o.y += __shfl_xor_sync(0xffffffffu, o.y, sh);
        o.z += __shfl_xor_sync(0xffffffffu, o.z, sh), o.w += __shfl_xor_sync(0xffffffffu, o.w, sh);
      }
      if (js == 0) *reinterpret_cast<float4*>(orow + 4 * lc) = o;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ backward ------------
// One CTA per (graph, head).  Rows are processed in rounds of 8 (one per warp): phase A builds
// dS_i and dQ_i for the warp's row, phase B lets every thread fold the round's 8 rows into the
// (j, c) entries of dK / dV it owns -- deterministic, no atomics.
template <int DH, int NCH>
__global__ void __launch_bounds__(kAttnThreads) attn_bwd_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t sn, int64_t sb,
    const uint8_t* __restrict__ mask, const float* __restrict__ attn, const float* __restrict__ rowflag,
    const float* __restrict__ d_o, int64_t osn, int64_t osb, const float* __restrict__ d_attn,
    float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, int64_t dsn, int64_t dsb, int H, int nmax,
    float scale, const float* __restrict__ drop) {
  extern __shared__ float smem[];
  __shared__ int s_neff;
  __shared__ int s_valid[kAttnWarps];
  const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
  const uint8_t* mk = mask + (size_t)b * nmax;
  const int n = block_n_eff(mk, nmax, &s_neff);
  const int npad = nmax | 1;
  float* Vt = smem;                              // [DH][npad]
  float* Ks = Vt + (size_t)DH * npad;            // [nmax][DH]
  float* Qs = Ks + (size_t)nmax * DH;            // [nmax][DH]
  float* dOs = Qs + (size_t)nmax * DH;           // [nmax][DH]
  float* dKs = dOs + (size_t)nmax * DH;          // [nmax][DH]
  float* dVs = dKs + (size_t)nmax * DH;          // [nmax][DH]
  float* Pb = dVs + (size_t)nmax * DH;           // [warps][npad]
  float* dSb = Pb + (size_t)kAttnWarps * npad;   // [warps][npad]

  const int64_t base_in = (int64_t)b * sb + h * DH;
  for (int idx = threadIdx.x; idx < nmax * DH; idx += blockDim.x) {
    const int j = idx / DH, c = idx - j * DH;
    float kk = 0.f, vv = 0.f, qq = 0.f, dd = 0.f;
    if (j < n) {
      kk = __ldg(k + base_in + (int64_t)j * sn + c);
      vv = __ldg(v + base_in + (int64_t)j * sn + c);
      qq = __ldg(q + base_in + (int64_t)j * sn + c);
      dd = __ldg(d_o + (int64_t)j * osn + (int64_t)b * osb + h * DH + c);
    }
    Vt[c * npad + j] = vv;
    Ks[idx] = kk;
    Qs[idx] = qq;
    dOs[idx] = dd;
    dKs[idx] = 0.0f;
    dVs[idx] = 0.0f;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* pr = Pb + (size_t)warp * npad;
  float* ds = dSb + (size_t)warp * npad;
  for (int base = 0; base < n; base += kAttnWarps) {
    const int i = base + warp;
    const bool valid = (i < n) && (mk[i] == 0);
    if (lane == 0) s_valid[warp] = valid;
    if (valid) {
      // ---- phase A
      float dor[DH];
#pragma unroll
      for (int c = 0; c < DH; ++c) dor[c] = dOs[i * DH + c];
      const float* arow = attn + (((size_t)b * H + h) * nmax + i) * nmax;
      const float* garow = d_attn ? d_attn + (((size_t)b * H + h) * nmax + i) * nmax : nullptr;
      const float* drow = drop ? drop + (((size_t)b * H + h) * nmax + i) * nmax : nullptr;
      float p[NCH], dp[NCH], pp[NCH];      // P (pre-dropout), dP (pre-dropout), P * drop
      float delta = 0.0f;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int j = lane + 32 * ch;
        float pv = 0.0f, dpv = 0.0f, ppv = 0.0f;
        if (j < n) {
          pv = __ldg(arow + j);
#pragma unroll
          for (int c = 0; c < DH; ++c) dpv = fmaf(dor[c], Vt[c * npad + j], dpv);
          if (garow) dpv += __ldg(garow + j);
          ppv = pv;
          if (drow) {
            const float dm = __ldg(drow + j);
            dpv *= dm;
            ppv = pv * dm;
          }
        }
        p[ch] = pv;
        dp[ch] = dpv;
        pp[ch] = ppv;
        delta = fmaf(pv, dpv, delta);
      }
      // rowflag 1: dS = P (dP - delta); rowflag -(1 + j*): clamped row, dS = P dP - [j = j*] delta (see forward)
      const float rf = __ldg(rowflag + ((size_t)b * H + h) * nmax + i);
      const float dsum = warp_sum(delta);
      delta = rf == 1.0f ? dsum : 0.0f;
      const int jstar = rf < 0.0f ? (int)(-rf) - 1 : -1;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int j = lane + 32 * ch;
        if (j < n) {
          pr[j] = pp[ch];
          ds[j] = p[ch] * (dp[ch] - delta) - (j == jstar ? dsum : 0.0f);
        }
      }
      __syncwarp();
      float* dqrow = dq + (int64_t)i * dsn + (int64_t)b * dsb + h * DH;
      if (DH <= 32) {
        constexpr int NS = DH <= 32 ? 32 / DH : 1;
        const int c = lane % DH, js = lane / DH;
        float a = 0.0f;
        for (int j = js; j < n; j += NS) a = fmaf(ds[j], Ks[j * DH + c], a);
        a = slice_reduce<(DH <= 32 ? DH : 32)>(a);
        if (js == 0) dqrow[c] = a * scale;
      } else {
        float a0 = 0.0f, a1 = 0.0f;
        for (int j = 0; j < n; ++j) {
          a0 = fmaf(ds[j], Ks[j * DH + lane], a0);
          a1 = fmaf(ds[j], Ks[j * DH + lane + 32], a1);
        }
        dqrow[lane] = a0 * scale;
        dqrow[lane + 32] = a1 * scale;
      }
    }
    __syncthreads();
    // ---- phase B
    for (int idx = threadIdx.x; idx < n * DH; idx += blockDim.x) {
      const int j = idx / DH, c = idx - j * DH;
      float ak = dKs[idx], av = dVs[idx];
#pragma unroll
      for (int w = 0; w < kAttnWarps; ++w) {
        if (s_valid[w]) {
          const int iw = base + w;
          ak = fmaf(dSb[w * npad + j], Qs[iw * DH + c], ak);
          av = fmaf(Pb[w * npad + j], dOs[iw * DH + c], av);
        }
      }
      dKs[idx] = ak;
      dVs[idx] = av;
    }
    __syncthreads();
  }
  // write dK, dV for every position (zeros for padding) and dQ = 0 for padded queries
  for (int idx = threadIdx.x; idx < nmax * DH; idx += blockDim.x) {
    const int j = idx / DH, c = idx - j * DH;
    const int64_t o = (int64_t)j * dsn + (int64_t)b * dsb + h * DH + c;
    const bool real = (j < n) && (mk[j] == 0);
    dk[o] = real ? dKs[idx] * scale : 0.0f;
    dv[o] = real ? dVs[idx] : 0.0f;
    if (!real) dq[o] = 0.0f;
  }
}

// ------------------------------------------------------------------ backward, register-tiled (dh 8 / 16) --
// Same contract and round structure as attn_bwd_kernel, restructured around its instruction mix (1 shared
// load per FMA there): Q/K/V/dO rows sit row-major with stride DH+4 (conflict-free 128-bit row loads), a
// lane owns key j and reads V_j / K_j as float4s, dQ_i is a thread-tiled reduction over key slices, and every
// thread keeps ITS (key, 4-channel) entries of dK / dV in registers for the whole kernel -- per round of
// RND rows a thread does 24 FMAs per 8 shared loads instead of 2 per 4.  Deterministic (fixed order).
template <int DH, int NCH, int THREADS>
__global__ void __launch_bounds__(THREADS) attn_bwd_tiled_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t sn, int64_t sb,
    const uint8_t* __restrict__ mask, const float* __restrict__ attn, const float* __restrict__ rowflag,
    const float* __restrict__ d_o, int64_t osn, int64_t osb, const float* __restrict__ d_attn,
    float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, int64_t dsn, int64_t dsb, int H, int nmax,
    float scale, const float* __restrict__ drop) {
  constexpr int LD = DH + 4, C4 = DH / 4, JT = THREADS / C4, EMAX = (32 * NCH + JT - 1) / JT;
  constexpr int RND = THREADS / 32;   // rows per round (one per warp)
  extern __shared__ float smem[];
  __shared__ int s_neff;
  __shared__ int s_valid[RND];
  const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
  const uint8_t* mk = mask + (size_t)b * nmax;
  const int n = block_n_eff(mk, nmax, &s_neff);
  const int npad = nmax | 1;
  float* Ks = smem;                               // [nmax][LD]
  float* Vs = Ks + (size_t)nmax * LD;
  float* Qs = Vs + (size_t)nmax * LD;
  float* dOs = Qs + (size_t)nmax * LD;
  float* Pb = dOs + (size_t)nmax * LD;            // [RND][npad]
  float* dSb = Pb + (size_t)RND * npad;           // [RND][npad]

  const int64_t base_in = (int64_t)b * sb + h * DH;
  for (int idx = threadIdx.x; idx < nmax * C4; idx += blockDim.x) {
    const int j = idx / C4, c = (idx - j * C4) * 4;
    float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk, qq = kk, dd = kk;
    if (j < n) {
      kk = __ldg(reinterpret_cast<const float4*>(k + base_in + (int64_t)j * sn + c));
      vv = __ldg(reinterpret_cast<const float4*>(v + base_in + (int64_t)j * sn + c));
      qq = __ldg(reinterpret_cast<const float4*>(q + base_in + (int64_t)j * sn + c));
      dd = __ldg(reinterpret_cast<const float4*>(d_o + (int64_t)j * osn + (int64_t)b * osb + h * DH + c));
    }
    *reinterpret_cast<float4*>(Ks + j * LD + c) = kk;
    *reinterpret_cast<float4*>(Vs + j * LD + c) = vv;
    *reinterpret_cast<float4*>(Qs + j * LD + c) = qq;
    *reinterpret_cast<float4*>(dOs + j * LD + c) = dd;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* pr = Pb + (size_t)warp * npad;
  float* ds = dSb + (size_t)warp * npad;
  // this thread's dK / dV entries: keys jt, jt + JT, ..., channels 4*c4 .. 4*c4+3
  const int c4 = threadIdx.x % C4, jt = threadIdx.x / C4;
  float4 accK[EMAX], accV[EMAX];
#pragma unroll
  for (int e = 0; e < EMAX; ++e) accK[e] = accV[e] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int base = 0; base < n; base += RND) {
    const int i = base + warp;
    const bool valid = (i < n) && (mk[i] == 0);
    if (lane == 0) s_valid[warp] = valid;
    if (valid) {
      // ---- phase A: dP, delta, dS for row i
      float dor[DH];
#pragma unroll
      for (int c = 0; c < C4; ++c) {
        const float4 t = *reinterpret_cast<const float4*>(dOs + i * LD + 4 * c);
        dor[4 * c] = t.x, dor[4 * c + 1] = t.y, dor[4 * c + 2] = t.z, dor[4 * c + 3] = t.w;
      }
      const float* arow = attn + (((size_t)b * H + h) * nmax + i) * nmax;
      const float* garow = d_attn ? d_attn + (((size_t)b * H + h) * nmax + i) * nmax : nullptr;
      float p[NCH], dp[NCH], pp[NCH];
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int j = lane + 32 * ch;
        pp[ch] = 0.0f;
        p[ch] = j < n ? __ldg(arow + j) : 0.0f;
        dp[ch] = (garow && j < n) ? __ldg(garow + j) : 0.0f;
      }
      float delta = 0.0f;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int j = lane + 32 * ch;
        if (j < n) {
          float a = dp[ch];
#pragma unroll
          for (int c = 0; c < C4; ++c) {
            const float4 t = *reinterpret_cast<const float4*>(Vs + j * LD + 4 * c);
            a = fmaf(dor[4 * c], t.x, a), a = fmaf(dor[4 * c + 1], t.y, a);
            a = fmaf(dor[4 * c + 2], t.z, a), a = fmaf(dor[4 * c + 3], t.w, a);
          }
          float ppv = p[ch];
          if (drop != nullptr) {     // attention-weight dropout: dP_pre = dP_post * drop, dV uses P * drop
            const float dm = __ldg(drop + (((size_t)b * H + h) * nmax + i) * nmax + j);
            a *= dm;
            ppv *= dm;
          }
          dp[ch] = a;
          pp[ch] = ppv;
          delta = fmaf(p[ch], a, delta);
        }
      }
      // rowflag 1: dS = P (dP - delta); rowflag -(1 + j*): clamped row, dS = P dP - [j = j*] delta (see forward)
      const float rf = __ldg(rowflag + ((size_t)b * H + h) * nmax + i);
      const float dsum = warp_sum(delta);
      delta = rf == 1.0f ? dsum : 0.0f;
      const int jstar = rf < 0.0f ? (int)(-rf) - 1 : -1;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        const int j = lane + 32 * ch;
        if (j < n) {
          pr[j] = pp[ch];
          ds[j] = p[ch] * (dp[ch] - delta) - (j == jstar ? dsum : 0.0f);
        }
      }
      __syncwarp();
      // dQ_i = scale * sum_j dS[j] K[j]: lanes = (4-channel group, key slice)
      {
        constexpr int NS = 32 / C4;
        const int lc = lane % C4, js = lane / C4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = js; j < n; j += NS) {
          const float w = ds[j];
          const float4 t = *reinterpret_cast<const float4*>(Ks + j * LD + 4 * lc);
          a.x = fmaf(w, t.x, a.x), a.y = fmaf(w, t.y, a.y), a.z = fmaf(w, t.z, a.z), a.w = fmaf(w, t.w, a.w);
        }
#pragma unroll
        for (int o = 16; o >= C4; o >>= 1) {
          a.x += __shfl_xor_sync(0xffffffffu, a.x, o), a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
          a.z += __shfl_xor_sync(0xffffffffu, a.z, o), a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
        }
        if (js == 0)
          *reinterpret_cast<float4*>(dq + (int64_t)i * dsn + (int64_t)b * dsb + h * DH + 4 * lc) =
              make_float4(a.x * scale, a.y * scale, a.z * scale, a.w * scale);
      }
    }
    __syncthreads();
    // ---- phase B: fold the round's rows into this thread's dK / dV entries
#pragma unroll
    for (int w = 0; w < RND; ++w) {
      if (s_valid[w]) {
        const int iw = base + w;
        const float4 qv = *reinterpret_cast<const float4*>(Qs + iw * LD + 4 * c4);
        const float4 gv = *reinterpret_cast<const float4*>(dOs + iw * LD + 4 * c4);
#pragma unroll
        for (int e = 0; e < EMAX; ++e) {
          const int j = jt + JT * e;
          if (j < n) {
            const float sj = dSb[w * npad + j], pj = Pb[w * npad + j];
            accK[e].x = fmaf(sj, qv.x, accK[e].x), accK[e].y = fmaf(sj, qv.y, accK[e].y);
            accK[e].z = fmaf(sj, qv.z, accK[e].z), accK[e].w = fmaf(sj, qv.w, accK[e].w);
            accV[e].x = fmaf(pj, gv.x, accV[e].x), accV[e].y = fmaf(pj, gv.y, accV[e].y);
            accV[e].z = fmaf(pj, gv.z, accV[e].z), accV[e].w = fmaf(pj, gv.w, accV[e].w);
          }
        }
      }
    }
    __syncthreads();
  }
  // write dK, dV for every position (zeros for padding) and dQ = 0 for padded queries
#pragma unroll
  for (int e = 0; e < EMAX; ++e) {
    const int j = jt + JT * e;
    if (j < nmax) {
      const int64_t o = (int64_t)j * dsn + (int64_t)b * dsb + h * DH + 4 * c4;
      const bool real = (j < n) && (mk[j] == 0);
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(dk + o) =
          real ? make_float4(accK[e].x * scale, accK[e].y * scale, accK[e].z * scale, accK[e].w * scale) : z;
      *reinterpret_cast<float4*>(dv + o) = real ? accV[e] : z;
      if (!real) *reinterpret_cast<float4*>(dq + o) = z;
    }
  }
}

static size_t attn_fwd_smem(int dh, int nmax) {
  const int npad = nmax | 1;
  return ((size_t)dh * npad + (size_t)nmax * dh + (size_t)kAttnWarps * npad + nmax) * sizeof(float);
}
static size_t attn_bwd_smem(int dh, int nmax) {
  const int npad = nmax | 1;
  return ((size_t)dh * npad + 5 * (size_t)nmax * dh + 2 * (size_t)kAttnWarps * npad) * sizeof(float);
}

static size_t attn_fwd_tiled_smem(int dh, int nmax) {
  const int npad = nmax | 1;
  return (2 * (size_t)nmax * (dh + 4) + (size_t)kAttnWarps * npad + nmax) * sizeof(float);
}

template <int DH, int NCH>
static int launch_attn_fwd(const float* q, const float* k, const float* v, int64_t sn, int64_t sb, const float* pe,
                           const uint8_t* mask, float* attn, float* o_heads, int64_t osn, int64_t osb, float* rowflag,
                           int B, int H, int nmax, float scale, const float* drop, float* attn_post, cudaStream_t st) {
  if constexpr ((DH == 8 || DH == 16) && NCH >= 4 && NCH <= 8) {
    const uintptr_t ptrs = (uintptr_t)k | (uintptr_t)v | (uintptr_t)o_heads;
    if ((ptrs % 16) == 0 && ((sn | sb | osn | osb) % 4) == 0 && getenv("FETA_ATTN_FWD_LEGACY") == nullptr) {
      const size_t smem_t = attn_fwd_tiled_smem(DH, nmax);
      FETA_CUDA(cudaFuncSetAttribute(attn_fwd_tiled_kernel<DH, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)smem_t));
      dim3 grid_t((unsigned)ceil_div(nmax, kRowsPerCta), (unsigned)(B * H));
      attn_fwd_tiled_kernel<DH, NCH><<<grid_t, kAttnThreads, smem_t, st>>>(q, k, v, sn, sb, pe, mask, attn, o_heads, osn,
                                                                           osb, rowflag, H, nmax, scale, kRowsPerCta,
                                                                           drop, attn_post);
      FETA_LAUNCH_CHECK();
      return FETA_OK;
    }
  }
  const size_t smem = attn_fwd_smem(DH, nmax);
  FETA_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<DH, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // small graphs: one CTA per (graph, head) so K/V are staged once; larger ones: 32-row tiles
  const int rows_per_cta = nmax <= 64 ? ((nmax + kAttnWarps - 1) / kAttnWarps) * kAttnWarps : kRowsPerCta;
  dim3 grid((unsigned)ceil_div(nmax, rows_per_cta), (unsigned)(B * H));
  attn_fwd_kernel<DH, NCH><<<grid, kAttnThreads, smem, st>>>(q, k, v, sn, sb, pe, mask, attn, o_heads, osn, osb,
                                                             rowflag, H, nmax, scale, rows_per_cta, drop, attn_post);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}
constexpr int kAttnBwdTiledThreads = 512;   // 16 rows per round: half the barriers, twice the warps per SM
static size_t attn_bwd_tiled_smem(int dh, int nmax) {
  const int npad = nmax | 1;
  return (4 * (size_t)nmax * (dh + 4) + 2 * (size_t)(kAttnBwdTiledThreads / 32) * npad) * sizeof(float);
}
// the register-tiled kernel needs float4-aligned per-head slices and fits dh in {8, 16}, nmax <= 256
static bool attn_bwd_tiled_ok(int dh, int nmax, const void* q, const void* k, const void* v, const void* d_o,
                              const void* dq, const void* dk, const void* dv, int64_t sn, int64_t sb, int64_t osn,
                              int64_t osb, int64_t dsn, int64_t dsb) {
  // graphs of <= 64 nodes: the one-LDS-per-FMA kernel is as fast (ZINC shape 17.1 vs 18.4 us) -- latency bound
  if (!(dh == 8 || dh == 16) || nmax <= 64 || nmax > 256 || getenv("FETA_ATTN_BWD_LEGACY") != nullptr) return false;
  const uintptr_t ptrs = (uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)d_o | (uintptr_t)dq | (uintptr_t)dk |
                         (uintptr_t)dv;
  const int64_t strides = sn | sb | osn | osb | dsn | dsb;
  return (ptrs % 16) == 0 && (strides % 4) == 0 && attn_bwd_tiled_smem(dh, nmax) <= 220 * 1024;
}

template <int DH, int NCH>
static int launch_attn_bwd(const float* q, const float* k, const float* v, int64_t sn, int64_t sb,
                           const uint8_t* mask, const float* attn, const float* rowflag, const float* d_o,
                           int64_t osn, int64_t osb, const float* d_attn, float* dq, float* dk, float* dv, int64_t dsn,
                           int64_t dsb, int B, int H, int nmax, float scale, const float* drop, cudaStream_t st) {
  if constexpr ((DH == 8 || DH == 16) && NCH >= 4 && NCH <= 8) {
    if (attn_bwd_tiled_ok(DH, nmax, q, k, v, d_o, dq, dk, dv, sn, sb, osn, osb, dsn, dsb)) {
      const size_t smem_t = attn_bwd_tiled_smem(DH, nmax);
      FETA_CUDA(cudaFuncSetAttribute(attn_bwd_tiled_kernel<DH, NCH, kAttnBwdTiledThreads>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
      attn_bwd_tiled_kernel<DH, NCH, kAttnBwdTiledThreads><<<(unsigned)(B * H), kAttnBwdTiledThreads, smem_t, st>>>(
          q, k, v, sn, sb, mask, attn, rowflag, d_o, osn, osb, d_attn, dq, dk, dv, dsn, dsb, H, nmax, scale, drop);
      FETA_LAUNCH_CHECK();
      return FETA_OK;
    }
  }
  const size_t smem = attn_bwd_smem(DH, nmax);
  FETA_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<DH, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attn_bwd_kernel<DH, NCH><<<(unsigned)(B * H), kAttnThreads, smem, st>>>(q, k, v, sn, sb, mask, attn, rowflag, d_o,
                                                                          osn, osb, d_attn, dq, dk, dv, dsn, dsb, H,
                                                                          nmax, scale, drop);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

#define FETA_ATTN_NCH(DH_, CALLNAME, ...)                                   \
  if (nmax <= 32) return CALLNAME<DH_, 1>(__VA_ARGS__);                     \
  if (nmax <= 64) return CALLNAME<DH_, 2>(__VA_ARGS__);                     \
  if (nmax <= 128) return CALLNAME<DH_, 4>(__VA_ARGS__);                    \
  if (nmax <= 256) return CALLNAME<DH_, 8>(__VA_ARGS__);                    \
  if (nmax <= 512) return CALLNAME<DH_, 16>(__VA_ARGS__);                   \
  return CALLNAME<DH_, 32>(__VA_ARGS__);

#define FETA_ATTN_DISPATCH(CALLNAME, ...)                                   \
  switch (dh) {                                                             \
    case 4: { FETA_ATTN_NCH(4, CALLNAME, __VA_ARGS__) }                     \
    case 8: { FETA_ATTN_NCH(8, CALLNAME, __VA_ARGS__) }                     \
    case 16: { FETA_ATTN_NCH(16, CALLNAME, __VA_ARGS__) }                   \
    case 32: { FETA_ATTN_NCH(32, CALLNAME, __VA_ARGS__) }                   \
    case 64: { FETA_ATTN_NCH(64, CALLNAME, __VA_ARGS__) }                   \
    default: break;                                                         \
  }

}  // namespace feta

using namespace feta;

static int attn_fwd_impl(const float* q, const float* k, const float* v, int64_t sn, int64_t sb, const float* pe,
                         const uint8_t* mask, float* attn, float* o_heads, int64_t osn, int64_t osb, float* rowflag,
                         int B, int H, int nmax, int dh, float scale, int use_tensor_cores, const float* drop,
                         float* attn_post, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE((drop == nullptr) == (attn_post == nullptr), "attn_fwd: drop and attn_post go together");
  if (drop != nullptr) use_tensor_cores = 0;     // the tcgen05 path has no dropout variant
  FETA_REQUIRE(B >= 0 && H >= 1 && nmax >= 0 && dh >= 1, "attn_fwd: bad sizes");
  if (B == 0 || nmax == 0) return FETA_OK;
  FETA_REQUIRE(q && k && v && mask && attn && o_heads && rowflag, "attn_fwd: NULL pointer argument");
  if (use_tensor_cores) {
    const int rc = attn_fwd_tc_try(q, k, v, sn, sb, pe, mask, attn, o_heads, osn, osb, rowflag, B, H, nmax, dh, scale,
                                   st);
    if (rc <= 0) return rc;
  }
  if (nmax > 1024 || attn_fwd_smem(dh, nmax) > 220 * 1024) {
    set_last_error("attn_fwd: nmax=%d dh=%d exceeds the shared-memory tile (nmax <= 1024, %zu B smem)", nmax, dh,
                   attn_fwd_smem(dh, nmax));
    return FETA_EUNSUPPORTED;
  }
  FETA_ATTN_DISPATCH(launch_attn_fwd, q, k, v, sn, sb, pe, mask, attn, o_heads, osn, osb, rowflag, B, H, nmax, scale,
                     drop, attn_post, st);
  set_last_error("attn_fwd: head dim %d not in {4,8,16,32,64}", dh);
  return FETA_EUNSUPPORTED;
}

extern "C" int feta_attn_fwd(const float* q, const float* k, const float* v, int64_t sn, int64_t sb, const float* pe,
                             const uint8_t* mask, float* attn, float* o_heads, int64_t osn, int64_t osb, float* rowflag,
                             int B, int H, int nmax, int dh, float scale, int use_tensor_cores, void* stream_) {
  return attn_fwd_impl(q, k, v, sn, sb, pe, mask, attn, o_heads, osn, osb, rowflag, B, H, nmax, dh, scale,
                       use_tensor_cores, nullptr, nullptr, stream_);
}

extern "C" int feta_attn_fwd_dropout(const float* q, const float* k, const float* v, int64_t sn, int64_t sb,
                                     const float* pe, const uint8_t* mask, const float* drop, float* attn,
                                     float* attn_post, float* o_heads, int64_t osn, int64_t osb, float* rowflag, int B,
                                     int H, int nmax, int dh, float scale, void* stream_) {
  FETA_REQUIRE(drop && attn_post, "attn_fwd_dropout: NULL drop / attn_post");
  return attn_fwd_impl(q, k, v, sn, sb, pe, mask, attn, o_heads, osn, osb, rowflag, B, H, nmax, dh, scale, 0, drop,
                       attn_post, stream_);
}

static int attn_bwd_impl(const float* q, const float* k, const float* v, int64_t sn, int64_t sb,
                         const uint8_t* mask, const float* attn, const float* rowflag, const float* d_o_heads,
                         int64_t osn, int64_t osb, const float* d_attn, float* dq, float* dk, float* dv,
                         int64_t dsn, int64_t dsb, int B, int H, int nmax, int dh, float scale, const float* drop,
                         void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(B >= 0 && H >= 1 && nmax >= 0 && dh >= 1, "attn_bwd: bad sizes");
  if (B == 0 || nmax == 0) return FETA_OK;
  FETA_REQUIRE(q && k && v && mask && attn && rowflag && d_o_heads && dq && dk && dv,
               "attn_bwd: NULL pointer argument");
  if (nmax > 1024 || attn_bwd_smem(dh, nmax) > 220 * 1024) {
    set_last_error("attn_bwd: nmax=%d dh=%d exceeds the shared-memory tile (%zu B smem)", nmax, dh,
                   attn_bwd_smem(dh, nmax));
    return FETA_EUNSUPPORTED;
  }
  FETA_ATTN_DISPATCH(launch_attn_bwd, q, k, v, sn, sb, mask, attn, rowflag, d_o_heads, osn, osb, d_attn, dq, dk, dv,
                     dsn, dsb, B, H, nmax, scale, drop, st);
  set_last_error("attn_bwd: head dim %d not in {4,8,16,32,64}", dh);
  return FETA_EUNSUPPORTED;
}

extern "C" int feta_attn_bwd(const float* q, const float* k, const float* v, int64_t sn, int64_t sb,
                             const uint8_t* mask, const float* attn, const float* rowflag, const float* d_o_heads,
                             int64_t osn, int64_t osb, const float* d_attn, float* dq, float* dk, float* dv,
                             int64_t dsn, int64_t dsb, int B, int H, int nmax, int dh, float scale, void* stream_) {
  return attn_bwd_impl(q, k, v, sn, sb, mask, attn, rowflag, d_o_heads, osn, osb, d_attn, dq, dk, dv, dsn, dsb, B, H,
                       nmax, dh, scale, nullptr, stream_);
}

extern "C" int feta_attn_bwd_dropout(const float* q, const float* k, const float* v, int64_t sn, int64_t sb,
                                     const uint8_t* mask, const float* attn, const float* rowflag, const float* drop,
                                     const float* d_o_heads, int64_t osn, int64_t osb, const float* d_attn_post,
                                     float* dq, float* dk, float* dv, int64_t dsn, int64_t dsb, int B, int H, int nmax,
                                     int dh, float scale, void* stream_) {
  FETA_REQUIRE(drop != nullptr, "attn_bwd_dropout: NULL drop");
  return attn_bwd_impl(q, k, v, sn, sb, mask, attn, rowflag, d_o_heads, osn, osb, d_attn_post, dq, dk, dv, dsn, dsb, B,
                       H, nmax, dh, scale, drop, stream_);
}
