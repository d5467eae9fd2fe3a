// Shared device/host helpers of the chunk-of-graphs kernels (cheb.cu, arma.cu): a CTA owns a chunk of
// consecutive whole graphs (block-diagonal operator => the chunk is closed under neighbours), one thread
// per row, rows staged in shared memory with a padded stride.
#pragma once
#include "common.cuh"

namespace feta {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// First row of the first graph whose start row is >= p (chunk boundaries; no search needed
// because row_graph gives the graph of row p).
__device__ __forceinline__ int chunk_boundary(int64_t p, int64_t R, const int32_t* __restrict__ graph_ptr,
                                              const int32_t* __restrict__ row_graph) {
  if (p >= R) return (int)R;
  const int g = row_graph[p];
  const int s = graph_ptr[g];
  return (s == (int)p) ? (int)p : graph_ptr[g + 1];
}

// device-side plan guard (see feta_cheb_fwd in include/feta_b200.h)
__device__ __forceinline__ bool plan_guard_ok(int32_t* meta, int64_t G, int max_nodes) {
  if (meta == nullptr) return true;
  const bool ok = meta[FETA_META_SORTED] == 1 && meta[FETA_META_BLOCKDIAG] == 1 &&
                  meta[FETA_META_NUM_GRAPHS] == (int32_t)G && meta[FETA_META_MAX_NODES] <= max_nodes &&
                  meta[FETA_META_BAD_INDEX] == 0;
  if (!ok && threadIdx.x == 0) meta[FETA_META_GUARD] = 1;
  return ok;
}

// A refused launch must not leave its (torch.empty) outputs as uninitialised memory that flows into the loss:
// every output is filled with NaN so the failure is visible downstream without a device->host read.
__device__ __forceinline__ void nan_fill(float* __restrict__ p, int64_t n) {
  if (p == nullptr) return;
  const float qnan = __int_as_float(0x7fc00000);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = qnan;
}
__device__ __forceinline__ void nan_fill_theta(float* __restrict__ p, int64_t sk, int64_t sg, int K, int64_t G,
                                               int FF) {
  if (p == nullptr) return;
  const float qnan = __int_as_float(0x7fc00000);
  const int64_t n = (int64_t)K * G * FF;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t kg = i / FF;
    p[(kg / G) * sk + (kg % G) * sg + (i - kg * FF)] = qnan;
  }
}

// coalesced copy of a [n, F] row-major global slab into shared memory with row stride LD
template <int F>
__device__ __forceinline__ void slab_to_smem(float* __restrict__ dst, const float* __restrict__ src, int n) {
  constexpr int LD = F + 4, Q = F / 4;
  for (int i = threadIdx.x; i < n * Q; i += blockDim.x) {
    const int row = i / Q, q = i - row * Q;
    st4(dst + row * LD + 4 * q, ldg4(src + (size_t)i * 4));
  }
}
template <int F>
__device__ __forceinline__ void smem_to_slab(float* __restrict__ dst, const float* __restrict__ src, int n) {
  constexpr int LD = F + 4, Q = F / 4;
  for (int i = threadIdx.x; i < n * Q; i += blockDim.x) {
    const int row = i / Q, q = i - row * Q;
    st4(dst + (size_t)i * 4, ld4(src + row * LD + 4 * q));
  }
}

// t[F] = sum_e vals[e] * buf[colidx[e] - r0]
template <int F>
__device__ __forceinline__ void gather_row(float (&t)[F], const float* __restrict__ buf, int r0,
                                           const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                           int e0, int e1) {
  constexpr int LD = F + 4;
#pragma unroll
  for (int i = 0; i < F; ++i) t[i] = 0.0f;
  for (int e = e0; e < e1; ++e) {
    const int c = __ldg(colidx + e) - r0;
    const float w = __ldg(vals + e);
    const float* row = buf + c * LD;
#pragma unroll
    for (int q = 0; q < F / 4; ++q) {
      const float4 a = ld4(row + 4 * q);
      t[4 * q + 0] = fmaf(w, a.x, t[4 * q + 0]);
      t[4 * q + 1] = fmaf(w, a.y, t[4 * q + 1]);
      t[4 * q + 2] = fmaf(w, a.z, t[4 * q + 2]);
      t[4 * q + 3] = fmaf(w, a.w, t[4 * q + 3]);
    }
  }
}

// acc[j] += sum_i t[i] * th[i*F + j]     (th: one graph's Theta_k, row-major [F, F], global/L1)
template <int F>
__device__ __forceinline__ void apply_theta(float (&acc)[F], const float (&t)[F], const float* __restrict__ th) {
#pragma unroll
  for (int i = 0; i < F; ++i) {
#pragma unroll
    for (int q = 0; q < F / 4; ++q) {
      const float4 w = ldg4(th + i * F + 4 * q);
      acc[4 * q + 0] = fmaf(t[i], w.x, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(t[i], w.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(t[i], w.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(t[i], w.w, acc[4 * q + 3]);
    }
  }
}

// d[i] = sum_j dout[j] * th[i*F + j]     (transpose application for the backward)
template <int F>
__device__ __forceinline__ void apply_theta_t(float (&d)[F], const float (&g)[F], const float* __restrict__ th) {
#pragma unroll
  for (int i = 0; i < F; ++i) {
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < F / 4; ++q) {
      const float4 w = ldg4(th + i * F + 4 * q);
      s = fmaf(g[4 * q + 0], w.x, s);
      s = fmaf(g[4 * q + 1], w.y, s);
      s = fmaf(g[4 * q + 2], w.z, s);
      s = fmaf(g[4 * q + 3], w.w, s);
    }
    d[i] = s;
  }
}

template <int F>
__device__ __forceinline__ void load_row(float (&t)[F], const float* __restrict__ row) {
#pragma unroll
  for (int q = 0; q < F / 4; ++q) {
    const float4 a = ld4(row + 4 * q);
    t[4 * q + 0] = a.x, t[4 * q + 1] = a.y, t[4 * q + 2] = a.z, t[4 * q + 3] = a.w;
  }
}
template <int F>
__device__ __forceinline__ void store_row(float* __restrict__ row, const float (&t)[F]) {
#pragma unroll
  for (int q = 0; q < F / 4; ++q) st4(row + 4 * q, make_float4(t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]));
}

struct FusedCfg {
  bool ok;
  int C, cap, threads;
  size_t smem;
};

// chunk quantum C and capacity (rows a chunk can hold) for the fused kernels
static inline FusedCfg fused_config(int F, int max_nodes, int nbuf, bool with_partial) {
  FusedCfg c{false, 0, 0, 0, 0};
  if (!(F == 4 || F == 8 || F == 16 || F == 32)) return c;
  if (max_nodes < 1) max_nodes = 1;
  const int maxT = F <= 8 ? 1024 : 512;
  int C = (max_nodes + 31) / 32 * 32;
  if (C < 64) C = 64;
  if (C > 256) C = 256;
  while (C > 32 && C + max_nodes - 1 > maxT) C -= 32;
  int cap = (C + max_nodes - 1 + 31) / 32 * 32;
  if (cap > maxT) return c;
  size_t smem = (size_t)nbuf * cap * (F + 4) * sizeof(float);
  if (with_partial) {
    const int NG = F * F / 4;
    smem += (size_t)(cap > NG ? cap : NG) * 16;
  }
  if (smem > 220 * 1024) return c;
  c.ok = true, c.C = C, c.cap = cap, c.threads = cap, c.smem = smem;
  return c;
}

#define FETA_DISPATCH_F(F_, ...)                            \
  switch (F_) {                                             \
    case 4: { constexpr int FF = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int FF = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int FF = 16; __VA_ARGS__; } break; \
    case 32: { constexpr int FF = 32; __VA_ARGS__; } break; \
    default: break;                                         \
  }


}  // namespace feta
