// A1/A3 for LARGE graphs (65 .. 256 rows: PATTERN / CLUSTER, SBM degree ~50): one CTA per graph, edge-parallel.
//
// The chunk kernel (cheb.cu) gives every row ONE thread, which then walks its ~50 neighbours serially three times
// (profiles/r1_cheb_chunk_pattern_ncu.md: 64 us, 10 % issue-active; with the reference's un-tiled edge_index only
// head 0's graphs carry edges at all -- SURVEY F4 -- so 1/4 of the CTAs do all the work).  Here a graph's rows,
// its Theta block and every Chebyshev order T_0..T_{K-1} live in shared memory, and the propagation is
// warp-per-row with lane = (edge slot, 16-byte chunk): 32/Q edges of a row are gathered per trip (Q = F/4), one
// LDS.128 + two FFMA2 each, then folded over the edge slots with log2(32/Q) shuffles.  The filter application is
// a separate thread-per-(row, 4 outputs) phase over the staged T_k.
// Backward, same residency, ONE launch:  T_k recomputed;  dTheta_k = T_k^T dOut per (k, i, 4 o's) thread (no
// atomics: the CTA owns the graph);  P_k = dOut Theta_k^T overwrites T_k;  Clenshaw  G_k = P_k + c_k L^T G_{k+1} -
// G_{k+2}  in place over the source-grouped CSR;  dx = G_0.
#include <stdlib.h>

#include "common.cuh"
#include "graph_tile.cuh"

namespace feta {
namespace dense {

constexpr int kThreads = 1024;   // 32 warps: the propagation is latency bound (dependent LDS chains)

template <int F>
struct Geo {
  static constexpr int Q = F / 4;        // 16-byte chunks per row
  static constexpr int ES = 32 / Q;      // edge slots per warp
  static constexpr int LD = F + 4;       // padded row stride (floats): conflict-free 16-byte row accesses
};

// acc (this lane's chunk q) = sum over the row's edges of vals[e] * src[colidx[e] - r0][q];  complete in lanes es == 0
template <int F, bool STAGED>
__device__ __forceinline__ float4 gather_row_edges(const float* __restrict__ src, int r0, const int32_t* __restrict__ ci,
                                                   const float* __restrict__ cv, int e0, int e1, int es, int q) {
  constexpr int ES = Geo<F>::ES, LD = Geo<F>::LD, Q = Geo<F>::Q;
  float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
  // staged slices hold the neighbour's float offset (row * LD), not its id.  Main loop: four edge slots per lane
  // and trip, no predicates -- index / weight loads and the four row loads are issued before the first FMA.
  constexpr int U = 4;
  const float* srcq = src + 4 * q;
  int e = e0 + es;
  for (; e + (U - 1) * ES < e1; e += U * ES) {
    int c[U];
    float w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      c[u] = STAGED ? ci[e + u * ES] : (__ldg(ci + e + u * ES) - r0) * LD;
      w[u] = STAGED ? cv[e + u * ES] : __ldg(cv + e + u * ES);
    }
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld4(srcq + c[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float2 ww = make_float2(w[u], w[u]);
      a0 = __ffma2_rn(ww, make_float2(v[u].x, v[u].y), a0);
      a1 = __ffma2_rn(ww, make_float2(v[u].z, v[u].w), a1);
    }
  }
  for (; e < e1; e += ES) {
    const int c = STAGED ? ci[e] : (__ldg(ci + e) - r0) * LD;
    const float w = STAGED ? cv[e] : __ldg(cv + e);
    const float4 v = ld4(srcq + c);
    const float2 ww = make_float2(w, w);
    a0 = __ffma2_rn(ww, make_float2(v.x, v.y), a0);
    a1 = __ffma2_rn(ww, make_float2(v.z, v.w), a1);
  }
  float4 acc = make_float4(a0.x, a0.y, a1.x, a1.y);
#pragma unroll
  for (int off = Q; off < 32; off <<= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
  }
  return acc;
}

// The graph's CSR slice in shared memory: every dependent L2 round trip (rowptr -> colidx / vals -> row) of the
// propagation becomes a shared-memory access.  Layout: [rp: cap + 8 ints][ci: ecap ints][cv: ecap floats].
struct StagedCsr {
  const int32_t* rp;    // rp[row] = first staged edge of the row (relative to the slice)
  const int32_t* ci;
  const float* cv;
  bool staged;
  int e_base;           // global index of the slice's first edge
};
template <int LD>
__device__ __forceinline__ StagedCsr stage_csr(int32_t* region, int cap, int ecap, int r0, int n,
                                               const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                               const float* __restrict__ vals) {
  StagedCsr c;
  int32_t* rp = region;
  int32_t* ci = region + cap + 8;
  float* cv = reinterpret_cast<float*>(ci + ecap);
  const int e_lo = __ldg(rowptr + r0), e_hi = __ldg(rowptr + r0 + n);
  c.staged = (e_hi - e_lo) <= ecap;
  c.e_base = e_lo;
  for (int i = threadIdx.x; i <= n; i += blockDim.x) rp[i] = __ldg(rowptr + r0 + i) - e_lo;
  if (c.staged)
    for (int i = threadIdx.x; i < e_hi - e_lo; i += blockDim.x) {
      ci[i] = (__ldg(colidx + e_lo + i) - r0) * LD;     // float offset of the neighbour's row
      cv[i] = __ldg(vals + e_lo + i);
    }
  c.rp = rp; c.ci = ci; c.cv = cv;
  return c;
}

// acc = sum of the row's edges out of `src` (this lane's chunk); complete in lanes es == 0
template <int F>
__device__ __forceinline__ float4 gather_row(const StagedCsr& c, const float* __restrict__ src, int r0, int row,
                                             const int32_t* __restrict__ colidx, const float* __restrict__ vals, int es,
                                             int q) {
  const int e0 = c.rp[row], e1 = c.rp[row + 1];
  if (e1 <= e0) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (c.staged) return gather_row_edges<F, true>(src, r0, c.ci, c.cv, e0, e1, es, q);
  return gather_row_edges<F, false>(src, r0, colidx, vals, c.e_base + e0, c.e_base + e1, es, q);
}

// T[k] = c L T[k-1] - T[k-2] for k = 1 .. K-1 (T[0] staged); T buffers are [K][cap][LD]
template <int F>
__device__ __forceinline__ void cheb_orders(float* __restrict__ T, int cap, int n, int r0, int K, const StagedCsr& c,
                                            const int32_t* __restrict__ colidx, const float* __restrict__ vals) {
  constexpr int LD = Geo<F>::LD, Q = Geo<F>::Q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int q = lane % Q, es = lane / Q;
  const size_t buf = (size_t)cap * LD;
  for (int k = 1; k < K; ++k) {
    const float* src = T + (size_t)(k - 1) * buf;
    float* dst = T + (size_t)k * buf;
    for (int row = warp; row < n; row += nw) {
      float4 acc = gather_row<F>(c, src, r0, row, colidx, vals, es, q);
      if (es == 0) {
        if (k >= 2) {
          const float4 o = ld4(dst - 2 * buf + row * LD + 4 * q);
          acc = make_float4(fmaf(2.f, acc.x, -o.x), fmaf(2.f, acc.y, -o.y), fmaf(2.f, acc.z, -o.z),
                            fmaf(2.f, acc.w, -o.w));
        }
        st4(dst + row * LD + 4 * q, acc);
      }
    }
    __syncthreads();
  }
}

template <int F>
__device__ __forceinline__ void stage_rows(float* __restrict__ dst, const float* __restrict__ src, int n) {
  constexpr int LD = Geo<F>::LD, Q = Geo<F>::Q;
  for (int i = threadIdx.x; i < n * Q; i += blockDim.x) {
    const int row = i / Q, q = i - row * Q;
    st4(dst + row * LD + 4 * q, ldg4(src + (size_t)i * 4));
  }
}

// =====================================================================================================
template <int F>
__global__ void __launch_bounds__(kThreads) cheb_graph_fwd_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
    const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr, const float* __restrict__ theta, int64_t sk,
    int64_t sg, const float* __restrict__ bias, float* __restrict__ out, int64_t R, int K, int cap, int ecap,
    int32_t* meta, int64_t G, int max_nodes) {
  constexpr int LD = Geo<F>::LD, Q = Geo<F>::Q;
  extern __shared__ float4 smem_f4[];
  if (!plan_guard_ok(meta, G, max_nodes)) { nan_fill(out, R * F); return; }
  float* T = reinterpret_cast<float*>(smem_f4);                    // [K][cap][LD]
  float* Th = T + (size_t)K * cap * LD;                            // [K][F][F]
  int32_t* csr_region = reinterpret_cast<int32_t*>(Th + (size_t)K * F * F);
  const int g = blockIdx.x;
  const int r0 = __ldg(graph_ptr + g), n = __ldg(graph_ptr + g + 1) - r0;
  if (n <= 0) return;
  stage_rows<F>(T, x + (size_t)r0 * F, n);
  for (int i = threadIdx.x; i < K * F * Q; i += blockDim.x) {
    const int k = i / (F * Q), rem = i - k * (F * Q);
    st4(Th + (size_t)i * 4, ldg4(theta + (int64_t)k * sk + (int64_t)g * sg + (int64_t)rem * 4));
  }
  const StagedCsr csr = stage_csr<LD>(csr_region, cap, ecap, r0, n, rowptr, colidx, vals);
  __syncthreads();
  cheb_orders<F>(T, cap, n, r0, K, csr, colidx, vals);
  // filter application: thread = (row, 4 outputs)
  const size_t buf = (size_t)cap * LD;
  for (int idx = threadIdx.x; idx < n * Q; idx += blockDim.x) {
    const int row = idx / Q, oq = idx - row * Q;
    float2 a0, a1;
    {
      const float4 b = bias ? ldg4(bias + 4 * oq) : make_float4(0.f, 0.f, 0.f, 0.f);
      a0 = make_float2(b.x, b.y), a1 = make_float2(b.z, b.w);
    }
    for (int k = 0; k < K; ++k) {
      const float* trow = T + (size_t)k * buf + row * LD;
      const float* th = Th + (size_t)k * F * F + 4 * oq;
#pragma unroll
      for (int iq = 0; iq < Q; ++iq) {
        const float4 t = ld4(trow + 4 * iq);
        const float tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const float4 w = ld4(th + (4 * iq + ii) * F);
          const float2 t2 = make_float2(tt[ii], tt[ii]);
          a0 = __ffma2_rn(t2, make_float2(w.x, w.y), a0);
          a1 = __ffma2_rn(t2, make_float2(w.z, w.w), a1);
        }
      }
    }
    st4(out + (size_t)(r0 + row) * F + 4 * oq, make_float4(a0.x, a0.y, a1.x, a1.y));
  }
}

// =====================================================================================================
template <int F>
__global__ void __launch_bounds__(kThreads) cheb_graph_bwd_kernel(
    const float* __restrict__ dout, const float* __restrict__ x, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ colidx, const float* __restrict__ vals, const int32_t* __restrict__ rowptr_t,
    const int32_t* __restrict__ colidx_t, const float* __restrict__ vals_t, const int32_t* __restrict__ graph_ptr,
    const float* __restrict__ theta, int64_t sk, int64_t sg, float* __restrict__ dx, float* __restrict__ dtheta,
    int64_t R, int K, int cap, int ecap, int32_t* meta, int64_t G, int max_nodes) {
  constexpr int LD = Geo<F>::LD, Q = Geo<F>::Q;
  extern __shared__ float4 smem_f4[];
  if (!plan_guard_ok(meta, G, max_nodes)) {
    nan_fill(dx, R * F);
    nan_fill_theta(dtheta, sk, sg, K, G, F * F);
    return;
  }
  const size_t buf = (size_t)cap * LD;
  float* T = reinterpret_cast<float*>(smem_f4);                    // [K][cap][LD]: T_k, then P_k / G_k
  float* D = T + (size_t)K * buf;                                  // dOut rows
  float* Th = D + buf;                                             // [K][F][F]
  int32_t* csr_region = reinterpret_cast<int32_t*>(Th + (size_t)K * F * F);
  const int g = blockIdx.x;
  const int r0 = __ldg(graph_ptr + g), n = __ldg(graph_ptr + g + 1) - r0;
  if (n <= 0) {      // an empty graph still owns a dTheta block
    if (dtheta)
      for (int i = threadIdx.x; i < K * F * F; i += blockDim.x)
        dtheta[(int64_t)(i / (F * F)) * sk + (int64_t)g * sg + (i % (F * F))] = 0.0f;
    return;
  }
  stage_rows<F>(D, dout + (size_t)r0 * F, n);
  for (int i = threadIdx.x; i < K * F * Q; i += blockDim.x) {
    const int k = i / (F * Q), rem = i - k * (F * Q);
    st4(Th + (size_t)i * 4, ldg4(theta + (int64_t)k * sk + (int64_t)g * sg + (int64_t)rem * 4));
  }
  if (dtheta != nullptr) {
    stage_rows<F>(T, x + (size_t)r0 * F, n);
    const StagedCsr csr = stage_csr<LD>(csr_region, cap, ecap, r0, n, rowptr, colidx, vals);
    __syncthreads();
    cheb_orders<F>(T, cap, n, r0, K, csr, colidx, vals);
    // dTheta_k[i][4oq..] = sum_rows T_k[row][i] * dOut[row][4oq..]: thread = (k, i, oq), lanes vary oq fastest
    for (int idx = threadIdx.x; idx < K * F * Q; idx += blockDim.x) {
      const int k = idx / (F * Q), rem = idx - k * (F * Q), i = rem / Q, oq = rem - i * Q;
      const float* tcol = T + (size_t)k * buf + i;
      const float* dcol = D + 4 * oq;
      float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
      for (int row = 0; row < n; ++row) {
        const float t = tcol[row * LD];
        const float4 d = ld4(dcol + row * LD);
        const float2 t2 = make_float2(t, t);
        a0 = __ffma2_rn(t2, make_float2(d.x, d.y), a0);
        a1 = __ffma2_rn(t2, make_float2(d.z, d.w), a1);
      }
      st4(dtheta + (int64_t)k * sk + (int64_t)g * sg + (int64_t)rem * 4, make_float4(a0.x, a0.y, a1.x, a1.y));
    }
  }
  __syncthreads();
  if (dx == nullptr) return;
  // P_k[row][4iq..] = sum_o dOut[row][o] * Theta_k[i][o]   (overwrites T_k)
  for (int idx = threadIdx.x; idx < K * n * Q; idx += blockDim.x) {
    const int k = idx / (n * Q), rem = idx - k * (n * Q), row = rem / Q, iq = rem - row * Q;
    const float* drow = D + row * LD;
    const float* th = Th + (size_t)k * F * F + (4 * iq) * F;
    float p[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int oq = 0; oq < Q; ++oq) {
      const float4 d = ld4(drow + 4 * oq);
#pragma unroll
      for (int ii = 0; ii < 4; ++ii) {
        const float4 w = ld4(th + ii * F + 4 * oq);
        p[ii] = fmaf(d.x, w.x, fmaf(d.y, w.y, fmaf(d.z, w.z, fmaf(d.w, w.w, p[ii]))));
      }
    }
    st4(T + (size_t)k * buf + row * LD + 4 * iq, make_float4(p[0], p[1], p[2], p[3]));
  }
  // the source-grouped CSR replaces the target-grouped one in the staging region (its readers are past the barrier)
  const StagedCsr csr_t = stage_csr<LD>(csr_region, cap, ecap, r0, n, rowptr_t, colidx_t, vals_t);
  __syncthreads();
  // Clenshaw, in place: G_k = P_k + c_k L^T G_{k+1} - G_{k+2}
  {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int q = lane % Q, es = lane / Q;
    for (int k = K - 2; k >= 0; --k) {
      float* gk = T + (size_t)k * buf;
      const float* g1 = gk + buf;
      const float ck = k == 0 ? 1.0f : 2.0f;
      for (int row = warp; row < n; row += nw) {
        const float4 acc = gather_row<F>(csr_t, g1, r0, row, colidx_t, vals_t, es, q);
        if (es == 0) {
          float4 p = ld4(gk + row * LD + 4 * q);
          p = make_float4(fmaf(ck, acc.x, p.x), fmaf(ck, acc.y, p.y), fmaf(ck, acc.z, p.z), fmaf(ck, acc.w, p.w));
          if (k + 2 < K) {
            const float4 o = ld4(gk + 2 * buf + row * LD + 4 * q);
            p = make_float4(p.x - o.x, p.y - o.y, p.z - o.z, p.w - o.w);
          }
          st4(gk + row * LD + 4 * q, p);
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n * Q; i += blockDim.x) {
    const int row = i / Q, q = i - row * Q;
    st4(dx + (size_t)r0 * F + (size_t)i * 4, ld4(T + row * LD + 4 * q));
  }
}

static inline int row_cap(int max_nodes) { return (max_nodes + 7) / 8 * 8; }
// staged edges per graph: a quarter of all node pairs (SBM density 0.35 .. 0.5 over ~1/2 of the pairs) or what fits
static inline int edge_cap(size_t base_bytes, int cap) {
  size_t room = (220 * 1024 - base_bytes) / 8;
  size_t want = (size_t)cap * cap / 2;
  size_t e = room < want ? room : want;
  return (int)(e / 4 * 4);
}

template <int F>
static int launch_fwd(const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                      const int32_t* graph_ptr, const float* theta, int64_t sk, int64_t sg, const float* bias,
                      float* out, int64_t R, int64_t G, int K, int32_t* meta, int max_nodes, cudaStream_t st) {
  const int cap = row_cap(max_nodes);
  const size_t base = ((size_t)K * cap * Geo<F>::LD + (size_t)K * F * F + cap + 8) * 4;
  if (base > 160 * 1024) return 1;
  const int ecap = edge_cap(base, cap);
  const size_t smem = base + (size_t)ecap * 8;
  FETA_CUDA(cudaFuncSetAttribute(cheb_graph_fwd_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cheb_graph_fwd_kernel<F><<<(unsigned)G, kThreads, smem, st>>>(x, rowptr, colidx, vals, graph_ptr, theta, sk, sg, bias,
                                                               out, R, K, cap, ecap, meta, G, max_nodes);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

template <int F>
static int launch_bwd(const float* dout, const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                      const int32_t* rowptr_t, const int32_t* colidx_t, const float* vals_t, const int32_t* graph_ptr,
                      const float* theta, int64_t sk, int64_t sg, float* dx, float* dtheta, int64_t R, int64_t G, int K,
                      int32_t* meta, int max_nodes, cudaStream_t st) {
  const int cap = row_cap(max_nodes);
  const size_t base = ((size_t)(K + 1) * cap * Geo<F>::LD + (size_t)K * F * F + cap + 8) * 4;
  if (base > 160 * 1024) return 1;
  const int ecap = edge_cap(base, cap);
  const size_t smem = base + (size_t)ecap * 8;
  FETA_CUDA(cudaFuncSetAttribute(cheb_graph_bwd_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cheb_graph_bwd_kernel<F><<<(unsigned)G, kThreads, smem, st>>>(dout, x, rowptr, colidx, vals, rowptr_t, colidx_t, vals_t,
                                                               graph_ptr, theta, sk, sg, dx, dtheta, R, K, cap, ecap,
                                                               meta, G, max_nodes);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

}  // namespace dense

// Eligibility: graphs of up to 256 rows, on average large enough that a 512-thread CTA per graph is not mostly
// idle (PATTERN / CLUSTER; molecule batches stay with the warp / tile kernels).
static bool dense_eligible(int64_t R, int64_t G, int F, int max_nodes) {
  if (getenv("FETA_CHEB_NO_DENSE_KERNEL") != nullptr) return false;
  return (F == 8 || F == 16) && max_nodes > 64 && max_nodes <= 256 && R >= 48 * G;
}

int cheb_fwd_dense_try(const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                       const int32_t* graph_ptr, const float* theta, int64_t sk, int64_t sg, const float* bias,
                       float* out, int64_t R, int64_t G, int K, int F, int max_nodes, int32_t* meta, cudaStream_t st) {
  if (!dense_eligible(R, G, F, max_nodes)) return 1;
  if (F == 16) return dense::launch_fwd<16>(x, rowptr, colidx, vals, graph_ptr, theta, sk, sg, bias, out, R, G, K, meta, max_nodes, st);
  return dense::launch_fwd<8>(x, rowptr, colidx, vals, graph_ptr, theta, sk, sg, bias, out, R, G, K, meta, max_nodes, st);
}

int cheb_bwd_dense_try(const float* dout, const float* x, const int32_t* rowptr, const int32_t* colidx,
                       const float* vals, const int32_t* rowptr_t, const int32_t* colidx_t, const float* vals_t,
                       const int32_t* graph_ptr, const float* theta, int64_t sk, int64_t sg, float* dx, float* dtheta,
                       int64_t R, int64_t G, int K, int F, int max_nodes, int32_t* meta, cudaStream_t st) {
  if (!dense_eligible(R, G, F, max_nodes)) return 1;
  if (F == 16)
    return dense::launch_bwd<16>(dout, x, rowptr, colidx, vals, rowptr_t, colidx_t, vals_t, graph_ptr, theta, sk, sg, dx,
                                 dtheta, R, G, K, meta, max_nodes, st);
  return dense::launch_bwd<8>(dout, x, rowptr, colidx, vals, rowptr_t, colidx_t, vals_t, graph_ptr, theta, sk, sg, dx,
                              dtheta, R, G, K, meta, max_nodes, st);
}

}  // namespace feta
