// A6 / N1: kernel-biased attention core WITHOUT a materialised attention matrix (see include/feta_b200.h).
//
// transformer/models.py:166-173 consumes the attention matrix of a layer only where the filter coefficients are built
// from it (the last layer under `last_layer_filter`); every other layer needs O = P V alone.  These kernels are the
// path for those layers: the forward pass keeps four numbers per query row (row maximum, reciprocal of the clamped
// kernel-weighted sum, ...) and the backward pass recomputes P from Q, K and the position-encoding kernel --
// 4 H n^2 bytes written and read back per layer are gone, and with them the instructions that moved them.
//
// Thread mapping ("lane per row"): one CTA per (graph, head); a THREAD owns one query row (forward, dQ) or one key
// row (dK, dV) and walks a contiguous slice of the other axis.  K / V (and Q / dO in the backward pass) of the
// (graph, head) sit in shared memory and are read as warp-wide broadcasts, so a score costs its dh FMAs plus dh/4
// LDS.128 for a whole warp of rows and no reduction crosses lanes -- the row-per-warp kernels of attention.cu spend
// most of their ~290 warp instructions per query row on shuffles, the slice fold and the attention-row store.
// A warp whose 32 rows are all padding leaves right after the staging barrier.  (A variant that split the other axis
// over several warps per row block to raise the number of resident warps was measured and dropped: every extra warp
// pays the same few hundred fixed instructions and the extra barriers cost more than the parallelism returned --
// ZINC backward 21 -> 23 us, PATTERN 61 -> 69 us, molhiv shape 83 -> 127 us; packed FFMA2 changed nothing either.)
// The position-encoding kernel is staged per warp as a 32-row x 16-column tile (coalesced reads, pitch 17: a lane
// reads ITS row conflict-free); the dK/dV phase needs no transposition and keeps its 16 values in registers.
//
//   forward :  m_i = max_j s_ij,  e_ij = 2^(s_ij - m_i) pe_ij,  O_i = sum_j e_ij V_j / max(sum_j e_ij, 1e-6)
//              with s_ij = log2(e) * scale * q_i . k_j (the constant is folded into q once per row)
//   backward:  delta_i = dO_i . O_i (= sum_j P_ij dP_ij),  dS_ij = P_ij (dO_i . V_j - delta_i)
//              phase 1 (thread = query row i): dQ_i = scale sum_j dS_ij K_j
//              phase 2 (thread = key row j)  : dK_j = scale sum_i dS_ij Q_i,   dV_j = sum_i P_ij dO_i
//   rows under the clamp (sum <= 1e-6, P = e pe / 1e-6): the denominator is a constant, but the row maximum is not
//   (oracle/layers.py:61-66 does not detach it), so dS_ij = P_ij dP_ij - [j = argmax_i] delta_i: the forward pass
//   records the argmax of such rows and both phases apply the correction.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace feta {
namespace arows {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kCW = 16;                 // key columns per staged position-encoding tile
constexpr int kTP = kCW + 1;            // its pitch
constexpr int kTile = 32 * kTP;         // floats per warp
constexpr int kMaxN = 256;               // one thread per padded position

template <int DH>
__device__ __forceinline__ void ld_row(float (&r)[DH], const float* __restrict__ p) {
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * c);
    r[4 * c] = t.x, r[4 * c + 1] = t.y, r[4 * c + 2] = t.z, r[4 * c + 3] = t.w;
  }
}
template <int DH>
__device__ __forceinline__ void ldg_row(float (&r)[DH], const float* __restrict__ p) {
#pragma unroll
  for (int c = 0; c < DH / 4; ++c) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + c);
    r[4 * c] = t.x, r[4 * c + 1] = t.y, r[4 * c + 2] = t.z, r[4 * c + 3] = t.w;
  }
}
template <int DH>
__device__ __forceinline__ float dot(const float (&a)[DH], const float (&b)[DH]) {
  float s0 = 0.0f, s1 = 0.0f;           // two chains: halves the dependent-FMA latency of a score
#pragma unroll
  for (int c = 0; c < DH; c += 2) s0 = fmaf(a[c], b[c], s0), s1 = fmaf(a[c + 1], b[c + 1], s1);
  return s0 + s1;
}
template <int DH>
__device__ __forceinline__ void axpy(float (&y)[DH], float a, const float (&x)[DH]) {
#pragma unroll
  for (int c = 0; c < DH; ++c) y[c] = fmaf(a, x[c], y[c]);
}

// tile A (lane = query row): reg[k] = pe_b[row0 + 2k + lane/16][col0 + lane%16]; rows / columns >= n read as 0
__device__ __forceinline__ void load_tile_a(float (&reg)[kCW], const float* __restrict__ peb, int nmax, int n, int row0,
                                            int col0, int lane) {
  const int c = col0 + (lane & 15), r = row0 + (lane >> 4);
  const float* src = peb + (size_t)r * nmax + c;
#pragma unroll
  for (int kk = 0; kk < kCW; ++kk)
    reg[kk] = (c < n && r + 2 * kk < n) ? __ldg(src + (size_t)(2 * kk) * nmax) : 0.0f;
}
__device__ __forceinline__ void store_tile_a(float* __restrict__ pes, const float (&reg)[kCW], int lane) {
  float* dst = pes + (lane >> 4) * kTP + (lane & 15);
#pragma unroll
  for (int kk = 0; kk < kCW; ++kk) dst[2 * kk * kTP] = reg[kk];
}
// tile B (lane = key column): reg[ii] = pe_b[i0 + ii][col]
__device__ __forceinline__ void load_tile_b(float (&reg)[kCW], const float* __restrict__ peb, int nmax, int n, int i0,
                                            int col) {
  const float* src = peb + (size_t)i0 * nmax + col;
#pragma unroll
  for (int ii = 0; ii < kCW; ++ii) reg[ii] = (col < n && i0 + ii < n) ? __ldg(src + (size_t)ii * nmax) : 0.0f;
}

// Padding masks are suffixes in every collate of the reference (data.py pads at the end): then n = number of real
// positions and the loops carry no per-key test.  Any other mask takes the GENERIC loops (n = nmax, per-key flag).
__device__ __forceinline__ int analyse_mask(bool valid, int t, int nmax, int* generic) {
  const int cnt = __syncthreads_count(valid);
  *generic = __syncthreads_or(valid && t >= cnt);
  return *generic ? nmax : cnt;
}

// ------------------------------------------------------------------ forward -------------
template <int DH, bool PE>
__global__ void __launch_bounds__(kMaxN) attn_rows_fwd_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t sn, int64_t sb,
    const float* __restrict__ pe, const uint8_t* __restrict__ mask, float* __restrict__ o_heads, int64_t osn,
    int64_t osb, float4* __restrict__ stats, int H, int nmax, float scale) {
  extern __shared__ __align__(16) float smem[];
  constexpr int C4 = DH / 4;
  const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int nr4 = (nmax + 3) & ~3;
  float* Ks = smem;                                  // [nmax][DH]
  float* Vs = Ks + (size_t)nmax * DH;                // [nmax][DH]
  float* vals = Vs + (size_t)nmax * DH;              // [nr4]  1 = real key
  float* pes = vals + nr4 + (size_t)warp * kTile;

  pdl_trigger();
  pdl_wait();          // q / k / v come from the previous kernel of the chain
  // every load that does not wait for the mask is issued before the mask barriers (one memory round trip)
  const int64_t hb = (int64_t)b * sb + h * DH;
  float qr[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) qr[c] = 0.0f;
  const bool valid = t < nmax && mask[(size_t)b * nmax + t] == 0;
  // the first warp loads its rows without waiting for the mask (it nearly always has real rows); the others load
  // only real rows -- molecule batches are padded to 3-4x their typical size
  if (t < nmax && (warp == 0 || valid)) {
    float kr[DH], vr[DH];
    ldg_row<DH>(kr, k + hb + (int64_t)t * sn);
    ldg_row<DH>(vr, v + hb + (int64_t)t * sn);
    ldg_row<DH>(qr, q + hb + (int64_t)t * sn);
#pragma unroll
    for (int c = 0; c < C4; ++c) {
      reinterpret_cast<float4*>(Ks + t * DH)[c] = make_float4(kr[4 * c], kr[4 * c + 1], kr[4 * c + 2], kr[4 * c + 3]);
      reinterpret_cast<float4*>(Vs + t * DH)[c] = make_float4(vr[4 * c], vr[4 * c + 1], vr[4 * c + 2], vr[4 * c + 3]);
    }
    const float f = scale * kLog2e;
#pragma unroll
    for (int c = 0; c < DH; ++c) qr[c] *= f;
  }
  if (t < nmax) vals[t] = valid ? 1.0f : 0.0f;
  int generic;
  const int n = analyse_mask(valid, t, nmax, &generic);          // its barriers publish Ks / Vs / vals

  float* orow = o_heads + (int64_t)t * osn + (int64_t)b * osb + h * DH;
  float4* srow = stats + ((size_t)b * H + h) * nmax + t;
  if (32 * warp >= n) {                              // no real row in this warp (no barrier follows)
    if (t < nmax) {
#pragma unroll
      for (int c = 0; c < C4; ++c) reinterpret_cast<float4*>(orow)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      *srow = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  const float* peb = PE ? pe + (size_t)b * nmax * nmax : nullptr;
  float per[kCW];
  if (PE) load_tile_a(per, peb, nmax, n, 32 * warp, 0, lane);    // in flight during pass 1

  float m = -INFINITY;
  if (!generic) {
#pragma unroll 4
    for (int j = 0; j < n; ++j) {
      float kj[DH];
      ld_row<DH>(kj, Ks + j * DH);
      m = fmaxf(m, dot<DH>(qr, kj));
    }
  } else {
    for (int j = 0; j < n; ++j) {
      if (vals[j] == 0.0f) continue;
      float kj[DH];
      ld_row<DH>(kj, Ks + j * DH);
      m = fmaxf(m, dot<DH>(qr, kj));
    }
  }

  float sum = 0.0f, o[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) o[c] = 0.0f;
  for (int jc = 0; jc < n; jc += kCW) {
    const int jn = min(kCW, n - jc);
    if (PE) {
      if (jc != 0) load_tile_a(per, peb, nmax, n, 32 * warp, jc, lane);
      store_tile_a(pes, per, lane);
      __syncwarp();
    }
    if (!generic) {
#pragma unroll 4
      for (int jj = 0; jj < jn; ++jj) {
        float kj[DH], vj[DH];
        ld_row<DH>(kj, Ks + (jc + jj) * DH);
        ld_row<DH>(vj, Vs + (jc + jj) * DH);
        float e = exp2f(dot<DH>(qr, kj) - m);
        if (PE) e *= pes[lane * kTP + jj];
        sum += e;
        axpy<DH>(o, e, vj);
      }
    } else {
      for (int jj = 0; jj < jn; ++jj) {
        if (vals[jc + jj] == 0.0f) continue;
        float kj[DH], vj[DH];
        ld_row<DH>(kj, Ks + (jc + jj) * DH);
        ld_row<DH>(vj, Vs + (jc + jj) * DH);
        float e = exp2f(dot<DH>(qr, kj) - m);
        if (PE) e *= pes[lane * kTP + jj];
        sum += e;
        axpy<DH>(o, e, vj);
      }
    }
    if (PE) __syncwarp();
  }

  if (t >= nmax) return;
  float4 st4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float inv = 0.0f;                                               // padded query: defined as zero
  if (t < n && valid) {
    inv = 1.0f / fmaxf(sum, 1e-6f);
    st4 = make_float4(m, inv, 0.0f, 1.0f);
    if (!(sum > 1e-6f)) {                                         // clamped row (rare): record the argmax
      float best = -INFINITY;
      int js = 0;
      for (int j = 0; j < n; ++j) {
        if (vals[j] == 0.0f) continue;
        float kj[DH];
        ld_row<DH>(kj, Ks + j * DH);
        const float s = dot<DH>(qr, kj);
        if (s > best) best = s, js = j;
      }
      st4.z = (float)js, st4.w = 2.0f;
    }
  }
  const bool rowok = st4.w != 0.0f;
#pragma unroll
  for (int c = 0; c < C4; ++c)
    reinterpret_cast<float4*>(orow)[c] =
        rowok ? make_float4(o[4 * c] * inv, o[4 * c + 1] * inv, o[4 * c + 2] * inv, o[4 * c + 3] * inv)
              : make_float4(0.f, 0.f, 0.f, 0.f);
  *srow = st4;
}

// ------------------------------------------------------------------ backward ------------
template <int DH, bool PE>
__global__ void __launch_bounds__(kMaxN) attn_rows_bwd_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t sn, int64_t sb,
    const float* __restrict__ pe, const uint8_t* __restrict__ mask, const float4* __restrict__ stats,
    const float* __restrict__ o_heads, const float* __restrict__ d_o, int64_t osn, int64_t osb,
    float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv, int64_t dsn, int64_t dsb, int H, int nmax,
    float scale) {
  extern __shared__ __align__(16) float smem[];
  constexpr int C4 = DH / 4;
  const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int nr4 = (nmax + 3) & ~3;
  float* Ks = smem;                                  // [nmax][DH]
  float* Vs = Ks + (size_t)nmax * DH;
  float* Qs = Vs + (size_t)nmax * DH;                // q * scale * log2(e)
  float* dOs = Qs + (size_t)nmax * DH;
  float4* sts = reinterpret_cast<float4*>(dOs + (size_t)nmax * DH);   // [nmax] (m, inv, delta, 0 | 1 | 2 + argmax)
  float* vals = reinterpret_cast<float*>(sts + nmax);                 // [nr4]
  float* pes = vals + nr4 + (size_t)warp * kTile;

  const int64_t hb = (int64_t)b * sb + h * DH, hob = (int64_t)b * osb + h * DH;
  pdl_trigger();
  pdl_wait();          // dO comes from the previous kernel of the chain
  float qr[DH], kr[DH], vr[DH], dor[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) qr[c] = kr[c] = vr[c] = dor[c] = 0.0f;
  float4 my = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool valid = t < nmax && mask[(size_t)b * nmax + t] == 0;
  if (t < nmax && (warp == 0 || valid)) {              // (see the forward kernel)
    ldg_row<DH>(kr, k + hb + (int64_t)t * sn);
    ldg_row<DH>(vr, v + hb + (int64_t)t * sn);
    ldg_row<DH>(qr, q + hb + (int64_t)t * sn);
    ldg_row<DH>(dor, d_o + hob + (int64_t)t * osn);
    float orow[DH];
    ldg_row<DH>(orow, o_heads + hob + (int64_t)t * osn);
    const float4 s4 = __ldg(stats + ((size_t)b * H + h) * nmax + t);
    const float f = scale * kLog2e;
#pragma unroll
    for (int c = 0; c < DH; ++c) qr[c] *= f;
    // delta_i = sum_j P_ij dP_ij = dO_i . O_i;  w: 0 padded, 1 normalised row, 2 + argmax for a clamped row
    my = make_float4(s4.x, s4.y, dot<DH>(dor, orow), s4.w == 2.0f ? 2.0f + s4.z : s4.w);
#pragma unroll
    for (int c = 0; c < C4; ++c) {
      reinterpret_cast<float4*>(Ks + t * DH)[c] = make_float4(kr[4 * c], kr[4 * c + 1], kr[4 * c + 2], kr[4 * c + 3]);
      reinterpret_cast<float4*>(Vs + t * DH)[c] = make_float4(vr[4 * c], vr[4 * c + 1], vr[4 * c + 2], vr[4 * c + 3]);
      reinterpret_cast<float4*>(Qs + t * DH)[c] = make_float4(qr[4 * c], qr[4 * c + 1], qr[4 * c + 2], qr[4 * c + 3]);
      reinterpret_cast<float4*>(dOs + t * DH)[c] =
          make_float4(dor[4 * c], dor[4 * c + 1], dor[4 * c + 2], dor[4 * c + 3]);
    }
  }
  if (t < nmax) {
    sts[t] = my;
    vals[t] = valid ? 1.0f : 0.0f;
  }
  int generic;
  const int n = analyse_mask(valid, t, nmax, &generic);
  if (t >= n || !valid) my.w = 0.0f;

  const int64_t ad = (int64_t)t * dsn + (int64_t)b * dsb + h * DH;
  float acc[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) acc[c] = 0.0f;
  if (32 * warp >= n) {                               // no real row in this warp (no barrier follows)
    if (t < nmax) {
#pragma unroll
      for (int c = 0; c < C4; ++c) {
        reinterpret_cast<float4*>(dq + ad)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(dk + ad)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(dv + ad)[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    return;
  }
  const float* peb = PE ? pe + (size_t)b * nmax * nmax : nullptr;

  // ---- phase 1: thread = query row t -> dQ_t
  {
    const float m = my.x, inv = my.y, delta = my.w == 1.0f ? my.z : 0.0f;
    float per[kCW];
    for (int jc = 0; jc < n; jc += kCW) {
      const int jn = min(kCW, n - jc);
      if (PE) {
        load_tile_a(per, peb, nmax, n, 32 * warp, jc, lane);
        store_tile_a(pes, per, lane);
        __syncwarp();
      }
      if (!generic) {
#pragma unroll 4
        for (int jj = 0; jj < jn; ++jj) {
          float kj[DH], vj[DH];
          ld_row<DH>(kj, Ks + (jc + jj) * DH);
          ld_row<DH>(vj, Vs + (jc + jj) * DH);
          float pr = exp2f(dot<DH>(qr, kj) - m) * inv;
          if (PE) pr *= pes[lane * kTP + jj];
          axpy<DH>(acc, pr * (dot<DH>(dor, vj) - delta), kj);
        }
      } else {
        for (int jj = 0; jj < jn; ++jj) {
          if (vals[jc + jj] == 0.0f) continue;
          float kj[DH], vj[DH];
          ld_row<DH>(kj, Ks + (jc + jj) * DH);
          ld_row<DH>(vj, Vs + (jc + jj) * DH);
          float pr = exp2f(dot<DH>(qr, kj) - m) * inv;
          if (PE) pr *= pes[lane * kTP + jj];
          axpy<DH>(acc, pr * (dot<DH>(dor, vj) - delta), kj);
        }
      }
      if (PE) __syncwarp();
    }
    if (my.w >= 2.0f) {                               // clamped row: the row maximum carries -delta to its argmax
      float kj[DH];
      ld_row<DH>(kj, Ks + ((int)my.w - 2) * DH);
      axpy<DH>(acc, -my.z, kj);
    }
    if (t < nmax) {
      const bool rowok = my.w != 0.0f;                // selects, not multiplies: a dead lane may hold inf / NaN
#pragma unroll
      for (int c = 0; c < C4; ++c)
        reinterpret_cast<float4*>(dq + ad)[c] =
            rowok ? make_float4(acc[4 * c] * scale, acc[4 * c + 1] * scale, acc[4 * c + 2] * scale,
                                acc[4 * c + 3] * scale)
                  : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }

  // ---- phase 2: thread = key row t -> dK_t, dV_t
  float dva[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) acc[c] = dva[c] = 0.0f;
  for (int ic = 0; ic < n; ic += kCW) {
    float per[kCW];
    if (PE) load_tile_b(per, peb, nmax, n, ic, t);
#pragma unroll
    for (int ii = 0; ii < kCW; ++ii) {
      if (ic + ii >= n) break;
      const float4 s4 = sts[ic + ii];
      if (s4.w == 0.0f) continue;                     // padded / masked query row (uniform)
      float qi[DH], gi[DH];
      ld_row<DH>(qi, Qs + (ic + ii) * DH);
      ld_row<DH>(gi, dOs + (ic + ii) * DH);
      float pr = exp2f(dot<DH>(qi, kr) - s4.x) * s4.y;
      if (PE) pr *= per[ii];
      const float delta = s4.w == 1.0f ? s4.z : 0.0f;
      axpy<DH>(acc, pr * (dot<DH>(gi, vr) - delta), qi);
      axpy<DH>(dva, pr, gi);
      if (s4.w >= 2.0f && (int)s4.w - 2 == t) axpy<DH>(acc, -s4.z, qi);
    }
  }
  if (t < nmax) {
    const bool keyok = t < n && valid;
#pragma unroll
    for (int c = 0; c < C4; ++c) {                    // Qs carries scale * log2(e): scale q = Qs ln 2
      reinterpret_cast<float4*>(dk + ad)[c] =
          keyok ? make_float4(acc[4 * c] * kLn2, acc[4 * c + 1] * kLn2, acc[4 * c + 2] * kLn2, acc[4 * c + 3] * kLn2)
                : make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(dv + ad)[c] =
          keyok ? make_float4(dva[4 * c], dva[4 * c + 1], dva[4 * c + 2], dva[4 * c + 3])
                : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ------------------------------------------------------------------ coefficient scalar (N1) ------------
// The filter-coefficient path (transformer/models.py:240-283) runs a GCN over an all-ones feature matrix on the
// complete graph weighted by the (detached) attention matrix; with x == 1 its output at node j collapses to
// s_j * colsum(W) + b (csrc/coeff.cu) with
//   loop_j = a_jj != 0 ? a_jj : 1,  deg_j = sum_{i != j} a_ij + loop_j,  dis = deg^-1/2,
//   s_j = dis_j (sum_{i != j} dis_i a_ij + dis_j loop_j).
// Both sums run over a COLUMN of the attention matrix, i.e. over query rows for a fixed key: exactly the key-owning
// thread mapping of the backward pass above.  This kernel recomputes a_ij = 2^(s_ij - m_i) pe_ij inv_i from q, k,
// the position-encoding kernel and the forward pass's row statistics -- twice, once per sum -- so the layer that feeds
// the coefficients needs no attention matrix either (no gradient flows here: the reference detaches it, :282).
template <int DH, bool PE>
__global__ void __launch_bounds__(kMaxN) attn_rows_coeff_kernel(
    const float* __restrict__ q, const float* __restrict__ k, int64_t sn, int64_t sb, const float* __restrict__ pe,
    const uint8_t* __restrict__ mask, const float4* __restrict__ stats, const int32_t* __restrict__ node_ptr,
    float* __restrict__ s_out, int H, int nmax, float scale, int64_t N) {
  extern __shared__ __align__(16) float smem[];
  constexpr int C4 = DH / 4;
  const int bh = blockIdx.x, b = bh / H, h = bh - b * H;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  float* Qs = smem;                                  // [nmax][DH]  q * scale * log2(e)
  float2* sts = reinterpret_cast<float2*>(Qs + (size_t)nmax * DH);    // [nmax] (m, inv); inv = 0: not a real row
  float* dis = reinterpret_cast<float*>(sts + nmax);                  // [nmax]
  int* wcnt = reinterpret_cast<int*>(dis + nmax);                     // [8] real positions per warp
  const int64_t hb = (int64_t)b * sb + h * DH;
  pdl_trigger();
  pdl_wait();
  const bool valid = t < nmax && mask[(size_t)b * nmax + t] == 0;
  float kr[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) kr[c] = 0.0f;
  if (valid) {
    float qr[DH];
    ldg_row<DH>(kr, k + hb + (int64_t)t * sn);
    ldg_row<DH>(qr, q + hb + (int64_t)t * sn);
    const float f = scale * kLog2e;
#pragma unroll
    for (int c = 0; c < C4; ++c)
      reinterpret_cast<float4*>(Qs + t * DH)[c] =
          make_float4(qr[4 * c] * f, qr[4 * c + 1] * f, qr[4 * c + 2] * f, qr[4 * c + 3] * f);
    const float4 s4 = __ldg(stats + ((size_t)b * H + h) * nmax + t);
    sts[t] = make_float2(s4.x, s4.w != 0.0f ? s4.y : 0.0f);
  } else if (t < nmax) {
    sts[t] = make_float2(0.0f, 0.0f);
  }
  // packed index of a real position = number of real positions before it
  const unsigned bal = __ballot_sync(0xffffffffu, valid);
  if (lane == 0) wcnt[warp] = __popc(bal);
  int generic;
  const int n = analyse_mask(valid, t, nmax, &generic);          // barriers: Qs / sts / wcnt published
  int rank = __popc(bal & ((1u << lane) - 1u));
  for (int w = 0; w < warp; ++w) rank += wcnt[w];
  const bool live = 32 * warp < n;                               // (no early return: one more barrier follows)
  const float* peb = PE ? pe + (size_t)b * nmax * nmax : nullptr;

  float deg = 0.0f, pjj = 0.0f;
  if (live) {
    for (int ic = 0; ic < n; ic += kCW) {
      float per[kCW];
      if (PE) load_tile_b(per, peb, nmax, n, ic, t);
#pragma unroll
      for (int ii = 0; ii < kCW; ++ii) {
        if (ic + ii >= n) break;
        const float2 st = sts[ic + ii];
        if (st.y == 0.0f) continue;                   // padded / masked query row: its matrix row is zero
        float qi[DH];
        ld_row<DH>(qi, Qs + (ic + ii) * DH);
        float p = exp2f(dot<DH>(qi, kr) - st.x) * st.y;
        if (PE) p *= per[ii];
        if (ic + ii == t) pjj = p;
        else deg += p;
      }
    }
  }
  const float lw = pjj != 0.0f ? pjj : 1.0f;          // add_remaining_self_loops keeps an existing loop weight
  deg += lw;
  const float d = (valid && deg > 0.0f) ? 1.0f / sqrtf(deg) : 0.0f;
  if (t < nmax) dis[t] = d;
  __syncthreads();
  if (!live || !valid) return;
  float acc = 0.0f;
  for (int ic = 0; ic < n; ic += kCW) {
    float per[kCW];
    if (PE) load_tile_b(per, peb, nmax, n, ic, t);
#pragma unroll
    for (int ii = 0; ii < kCW; ++ii) {
      if (ic + ii >= n) break;
      const float2 st = sts[ic + ii];
      if (st.y == 0.0f || ic + ii == t) continue;
      float qi[DH];
      ld_row<DH>(qi, Qs + (ic + ii) * DH);
      float p = exp2f(dot<DH>(qi, kr) - st.x) * st.y;
      if (PE) p *= per[ii];
      acc = fmaf(dis[ic + ii], p, acc);
    }
  }
  s_out[(int64_t)h * N + __ldg(node_ptr + b) + rank] = d * (acc + d * lw);
}

static size_t fwd_smem(int dh, int nmax, bool has_pe) {
  const int nr4 = (nmax + 3) & ~3, warps = (nmax + 31) / 32;
  return (2 * (size_t)nmax * dh + nr4 + (has_pe ? (size_t)warps * kTile : 0)) * sizeof(float);
}
static size_t bwd_smem(int dh, int nmax, bool has_pe) {
  const int nr4 = (nmax + 3) & ~3, warps = (nmax + 31) / 32;
  return (4 * (size_t)nmax * dh + 4 * (size_t)nmax + nr4 + (has_pe ? (size_t)warps * kTile : 0)) * sizeof(float);
}

template <typename K>
static int grant_smem(K kernel, std::atomic<int>& granted, size_t smem) {
  if ((int)smem > granted.load(std::memory_order_relaxed)) {     // the opt-in limit only ever grows
    FETA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    granted.store((int)smem, std::memory_order_relaxed);
  }
  return FETA_OK;
}

template <int DH, bool PE>
static int launch_fwd(const float* q, const float* k, const float* v, int64_t sn, int64_t sb, const float* pe,
                      const uint8_t* mask, float* o_heads, int64_t osn, int64_t osb, float* stats, int B, int H,
                      int nmax, float scale, cudaStream_t st) {
  const size_t smem = fwd_smem(DH, nmax, PE);
  static std::atomic<int> granted{48 * 1024};
  const int rc = grant_smem(attn_rows_fwd_kernel<DH, PE>, granted, smem);
  if (rc != FETA_OK) return rc;
  FETA_CUDA(launch_chain(attn_rows_fwd_kernel<DH, PE>, dim3((unsigned)(B * H)), dim3(32 * ((nmax + 31) / 32)), smem, st,
                         q, k, v, sn, sb, pe, mask, o_heads, osn, osb, reinterpret_cast<float4*>(stats), H, nmax,
                         scale));
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

template <int DH, bool PE>
static int launch_bwd(const float* q, const float* k, const float* v, int64_t sn, int64_t sb, const float* pe,
                      const uint8_t* mask, const float* stats, const float* o_heads, const float* d_o, int64_t osn,
                      int64_t osb, float* dq, float* dk, float* dv, int64_t dsn, int64_t dsb, int B, int H, int nmax,
                      float scale, cudaStream_t st) {
  const size_t smem = bwd_smem(DH, nmax, PE);
  static std::atomic<int> granted{48 * 1024};
  const int rc = grant_smem(attn_rows_bwd_kernel<DH, PE>, granted, smem);
  if (rc != FETA_OK) return rc;
  FETA_CUDA(launch_chain(attn_rows_bwd_kernel<DH, PE>, dim3((unsigned)(B * H)), dim3(32 * ((nmax + 31) / 32)), smem, st,
                         q, k, v, sn, sb, pe, mask, reinterpret_cast<const float4*>(stats), o_heads, d_o, osn, osb, dq,
                         dk, dv, dsn, dsb, H, nmax, scale));
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

template <int DH, bool PE>
static int launch_coeff(const float* q, const float* k, int64_t sn, int64_t sb, const float* pe, const uint8_t* mask,
                        const float* stats, const int32_t* node_ptr, float* s_out, int B, int H, int nmax, float scale,
                        int64_t N, cudaStream_t st) {
  const size_t smem = ((size_t)nmax * DH + 3 * (size_t)nmax + 8) * sizeof(float);
  static std::atomic<int> granted{48 * 1024};
  const int rc = grant_smem(attn_rows_coeff_kernel<DH, PE>, granted, smem);
  if (rc != FETA_OK) return rc;
  FETA_CUDA(launch_chain(attn_rows_coeff_kernel<DH, PE>, dim3((unsigned)(B * H)), dim3(32 * ((nmax + 31) / 32)), smem, st,
                         q, k, sn, sb, pe, mask, reinterpret_cast<const float4*>(stats), node_ptr, s_out, H, nmax,
                         scale, N));
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

#define FETA_AROWS_DISPATCH(FN, ...)                                   \
  switch (dh) {                                                        \
    case 4: return pe ? FN<4, true>(__VA_ARGS__) : FN<4, false>(__VA_ARGS__);     \
    case 8: return pe ? FN<8, true>(__VA_ARGS__) : FN<8, false>(__VA_ARGS__);     \
    case 16: return pe ? FN<16, true>(__VA_ARGS__) : FN<16, false>(__VA_ARGS__);  \
    case 32: return pe ? FN<32, true>(__VA_ARGS__) : FN<32, false>(__VA_ARGS__);  \
    default: break;                                                    \
  }

static bool aligned16(std::initializer_list<const void*> ptrs, std::initializer_list<int64_t> strides) {
  uintptr_t a = 0;
  for (const void* p : ptrs) a |= (uintptr_t)p;
  int64_t s = 0;
  for (int64_t x : strides) s |= x;
  return (a % 16) == 0 && (s % 4) == 0;
}

}  // namespace arows
}  // namespace feta

using namespace feta;

extern "C" int feta_attn_rows_supported(int nmax, int dh) {
  return nmax >= 1 && nmax <= arows::kMaxN && (dh == 4 || dh == 8 || dh == 16 || dh == 32);
}

extern "C" int feta_attn_rows_fwd(const float* q, const float* k, const float* v, int64_t sn, int64_t sb,
                                  const float* pe, const uint8_t* mask, float* o_heads, int64_t osn, int64_t osb,
                                  float* stats, int B, int H, int nmax, int dh, float scale, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(B >= 0 && H >= 1 && nmax >= 0 && dh >= 1, "attn_rows_fwd: bad sizes");
  if (B == 0 || nmax == 0) return FETA_OK;
  FETA_REQUIRE(q && k && v && mask && o_heads && stats, "attn_rows_fwd: NULL pointer argument");
  if (!feta_attn_rows_supported(nmax, dh) || !arows::aligned16({q, k, v, o_heads, stats}, {sn, sb, osn, osb})) {
    set_last_error("attn_rows_fwd: needs nmax <= %d, dh in {4,8,16,32}, 16-byte aligned head slices (nmax=%d dh=%d)",
                   arows::kMaxN, nmax, dh);
    return FETA_EUNSUPPORTED;
  }
  FETA_AROWS_DISPATCH(arows::launch_fwd, q, k, v, sn, sb, pe, mask, o_heads, osn, osb, stats, B, H, nmax, scale, st);
  return FETA_EUNSUPPORTED;
}

extern "C" int feta_attn_rows_bwd(const float* q, const float* k, const float* v, int64_t sn, int64_t sb,
                                  const float* pe, const uint8_t* mask, const float* stats, const float* o_heads,
                                  const float* d_o_heads, int64_t osn, int64_t osb, float* dq, float* dk, float* dv,
                                  int64_t dsn, int64_t dsb, int B, int H, int nmax, int dh, float scale,
                                  void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(B >= 0 && H >= 1 && nmax >= 0 && dh >= 1, "attn_rows_bwd: bad sizes");
  if (B == 0 || nmax == 0) return FETA_OK;
  FETA_REQUIRE(q && k && v && mask && stats && o_heads && d_o_heads && dq && dk && dv,
               "attn_rows_bwd: NULL pointer argument");
  if (!feta_attn_rows_supported(nmax, dh) ||
      !arows::aligned16({q, k, v, o_heads, d_o_heads, stats, dq, dk, dv}, {sn, sb, osn, osb, dsn, dsb})) {
    set_last_error("attn_rows_bwd: needs nmax <= %d, dh in {4,8,16,32}, 16-byte aligned head slices (nmax=%d dh=%d)",
                   arows::kMaxN, nmax, dh);
    return FETA_EUNSUPPORTED;
  }
  FETA_AROWS_DISPATCH(arows::launch_bwd, q, k, v, sn, sb, pe, mask, stats, o_heads, d_o_heads, osn, osb, dq, dk, dv,
                      dsn, dsb, B, H, nmax, scale, st);
  return FETA_EUNSUPPORTED;
}

extern "C" int feta_attn_rows_coeff(const float* q, const float* k, int64_t sn, int64_t sb, const float* pe,
                                    const uint8_t* mask, const float* stats, const int32_t* node_ptr, float* s_out,
                                    int B, int H, int nmax, int dh, float scale, int64_t N, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(B >= 0 && H >= 1 && nmax >= 0 && dh >= 1 && N >= 0, "attn_rows_coeff: bad sizes");
  if (B == 0 || nmax == 0 || N == 0) return FETA_OK;
  FETA_REQUIRE(q && k && mask && stats && node_ptr && s_out, "attn_rows_coeff: NULL pointer argument");
  if (!feta_attn_rows_supported(nmax, dh) || !arows::aligned16({q, k, stats}, {sn, sb})) {
    set_last_error("attn_rows_coeff: needs nmax <= %d, dh in {4,8,16,32}, 16-byte aligned head slices (nmax=%d dh=%d)",
                   arows::kMaxN, nmax, dh);
    return FETA_EUNSUPPORTED;
  }
  FETA_AROWS_DISPATCH(arows::launch_coeff, q, k, sn, sb, pe, mask, stats, node_ptr, s_out, B, H, nmax, scale, N, st);
  return FETA_EUNSUPPORTED;
}

