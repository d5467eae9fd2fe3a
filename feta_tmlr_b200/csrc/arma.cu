// N4: fused ARMAConvDynamic forward / backward (see include/feta_b200.h).
//
// Replaces ARMAConvDynamic.forward (transformer/ChebNetDynamic.py:297-346) for the only configuration
// the reference instantiates (num_layers = 1, transformer/models.py:139).  The reference materialises a
// per-node weight [R, K, F, F] twice (:323, :337), runs 2 batched 1xF.FxF bmm's over R*K problems
// (_batch_multiply_coeff, :274-295) and one PyG propagate on a [K, R, F] tensor.  Because the propagation
// is linear,  A_hat (x (a W_k)) = a (A_hat x) W_k,  so ONE gather per row serves all K stacks:
//
//   out[r] = 1/K sum_k relu( a[g,k] (A_hat x)[r] W_k + b[g,k] x_root[r] V_k + bias_k )
//
// Same tiling as the Chebyshev kernels: a CTA owns a chunk of whole graphs, one thread per row, rows
// staged in shared memory; HBM traffic = x + out + (A_hat x saved for the backward) + CSR.
#include "common.cuh"
#include "graph_tile.cuh"

namespace feta {

// acc[j] += sum_i t[i] w[i*F + j]   /   d[i] = sum_j g[j] w[i*F + j]   with w in SHARED memory (the stack
// weights are the same for every row: staged once per CTA, read as broadcasts)
template <int F>
__device__ __forceinline__ void apply_w(float (&acc)[F], const float (&t)[F], const float* __restrict__ w) {
#pragma unroll
  for (int i = 0; i < F; ++i) {
#pragma unroll
    for (int q = 0; q < F / 4; ++q) {
      const float4 ww = ld4(w + i * F + 4 * q);
      acc[4 * q + 0] = fmaf(t[i], ww.x, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(t[i], ww.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(t[i], ww.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(t[i], ww.w, acc[4 * q + 3]);
    }
  }
}
template <int F>
__device__ __forceinline__ void apply_w_t(float (&d)[F], const float (&g)[F], const float* __restrict__ w) {
#pragma unroll
  for (int i = 0; i < F; ++i) {
    float s = 0.0f;
#pragma unroll
    for (int q = 0; q < F / 4; ++q) {
      const float4 ww = ld4(w + i * F + 4 * q);
      s = fmaf(g[4 * q + 0], ww.x, s);
      s = fmaf(g[4 * q + 1], ww.y, s);
      s = fmaf(g[4 * q + 2], ww.z, s);
      s = fmaf(g[4 * q + 3], ww.w, s);
    }
    d[i] = s;
  }
}
// W | V | bias  ->  shared memory ([K,F,F], [K,F,F], [K,F]; bias zero-filled when absent)
template <int F>
__device__ __forceinline__ void stage_weights(float* __restrict__ ws, const float* __restrict__ W,
                                              const float* __restrict__ V, const float* __restrict__ bias, int K) {
  const int nw = K * F * F;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) {
    ws[i] = __ldg(W + i);
    ws[nw + i] = __ldg(V + i);
  }
  for (int i = threadIdx.x; i < K * F; i += blockDim.x) ws[2 * nw + i] = bias ? __ldg(bias + i) : 0.0f;
}

template <int F>
__global__ void __launch_bounds__(512) arma_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ xr, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ colidx, const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr,
    const int32_t* __restrict__ row_graph, const float* __restrict__ coeff, const float* __restrict__ W,
    const float* __restrict__ V, const float* __restrict__ bias, float* __restrict__ out, float* __restrict__ prop,
    int64_t R, int K, int C, int cap, int32_t* meta, int64_t G, int max_nodes) {
  constexpr int LD = F + 4;
  extern __shared__ float4 smem_f4[];
  if (!plan_guard_ok(meta, G, max_nodes)) { nan_fill(out, R * F); nan_fill(prop, R * F); return; }
  float* buf0 = reinterpret_cast<float*>(smem_f4);
  float* buf1 = buf0 + (size_t)cap * LD;
  float* ws = buf1 + (size_t)cap * LD;
  const float* Ws = ws;
  const float* Vs = ws + K * F * F;
  const float* bs = ws + 2 * K * F * F;
  const int64_t p0 = (int64_t)blockIdx.x * C;
  const int r0 = chunk_boundary(p0, R, graph_ptr, row_graph);
  const int r1 = chunk_boundary(p0 + C, R, graph_ptr, row_graph);
  const int n = r1 - r0;
  if (n <= 0) return;
  slab_to_smem<F>(buf0, x + (size_t)r0 * F, n);
  stage_weights<F>(ws, W, V, bias, K);
  __syncthreads();

  const int lr = threadIdx.x;
  const bool active = lr < n;
  const int r = r0 + lr;
  float p[F], acc[F];
  if (active) {
    gather_row<F>(p, buf0, r0, colidx, vals, rowptr[r], rowptr[r + 1]);   // propagate, :332-333
    float xrow[F];
    if (xr == x) {
      load_row<F>(xrow, buf0 + lr * LD);
    } else {
#pragma unroll
      for (int q = 0; q < F / 4; ++q) {
        const float4 a = ldg4(xr + (size_t)r * F + 4 * q);
        xrow[4 * q] = a.x, xrow[4 * q + 1] = a.y, xrow[4 * q + 2] = a.z, xrow[4 * q + 3] = a.w;
      }
    }
    const float* cg = coeff + (int64_t)row_graph[r] * 2 * K;
#pragma unroll
    for (int j = 0; j < F; ++j) acc[j] = 0.0f;
    for (int k = 0; k < K; ++k) {
      float u[F], v[F];
#pragma unroll
      for (int j = 0; j < F; ++j) u[j] = v[j] = 0.0f;
      apply_w<F>(u, p, Ws + k * F * F);      // init_weight * filter_coeff_a, :323-324
      apply_w<F>(v, xrow, Vs + k * F * F);   // root_weight * filter_coeff_b, :337-338
      const float a = __ldg(cg + k), b = __ldg(cg + K + k);
#pragma unroll
      for (int j = 0; j < F; ++j) {
        const float z = fmaf(a, u[j], fmaf(b, v[j], bs[k * F + j]));   // + bias, :340-341
        acc[j] += fmaxf(z, 0.0f);                                                              // act, :343-344
      }
    }
    const float inv = 1.0f / (float)K;   // mean over the stack axis, :346
#pragma unroll
    for (int j = 0; j < F; ++j) acc[j] *= inv;
    store_row<F>(buf1 + lr * LD, acc);
  }
  __syncthreads();   // every gather has read buf0
  if (active && prop != nullptr) store_row<F>(buf0 + lr * LD, p);
  __syncthreads();
  smem_to_slab<F>(out + (size_t)r0 * F, buf1, n);
  if (prop != nullptr) smem_to_slab<F>(prop + (size_t)r0 * F, buf0, n);
}

// Backward.  Per row and stack: dz_k = [z_k > 0] dout / K (z_k recomputed from the saved A_hat x);
//   d(A_hat x) = sum_k a dz_k W_k^T,  dx_root = sum_k b dz_k V_k^T,  dx = A_hat^T d(A_hat x)  (in-chunk gather
//   over the SOURCE-grouped CSR);  dcoeff[g] = per-graph sums of (pW_k).dz_k and (x_root V_k).dz_k, reduced in a
//   fixed order (deterministic);  dz, a.dz, b.dz are written out for the weight-gradient GEMMs.
template <int F>
__global__ void __launch_bounds__(256) arma_bwd_kernel(
    const float* __restrict__ dout, const float* __restrict__ prop, const float* __restrict__ xr,
    const int32_t* __restrict__ rowptr_t, const int32_t* __restrict__ colidx_t, const float* __restrict__ vals_t,
    const int32_t* __restrict__ graph_ptr, const int32_t* __restrict__ row_graph, const float* __restrict__ coeff,
    const float* __restrict__ W, const float* __restrict__ V, const float* __restrict__ bias,
    float* __restrict__ dx, float* __restrict__ dx_root, float* __restrict__ dz, float* __restrict__ dza,
    float* __restrict__ dzb, float* __restrict__ dcoeff, int64_t R, int K, int C, int cap, int32_t* meta, int64_t G,
    int max_nodes) {
  constexpr int LD = F + 4;
  extern __shared__ float4 smem_f4[];
  if (!plan_guard_ok(meta, G, max_nodes)) {
    nan_fill(dx, R * F); nan_fill(dx_root, R * F); nan_fill(dcoeff, G * 2 * K);
    nan_fill(dz, R * K * F); nan_fill(dza, R * K * F); nan_fill(dzb, R * K * F);
    return;
  }
  float* buf0 = reinterpret_cast<float*>(smem_f4);
  float* part = buf0 + (size_t)cap * LD;   // [cap, 2K] per-row coefficient partials
  float* ws = part + (size_t)cap * 2 * K;
  const float* Ws = ws;
  const float* Vs = ws + K * F * F;
  const float* bs = ws + 2 * K * F * F;
  const int64_t p0 = (int64_t)blockIdx.x * C;
  const int r0 = chunk_boundary(p0, R, graph_ptr, row_graph);
  const int r1 = chunk_boundary(p0 + C, R, graph_ptr, row_graph);
  const int n = r1 - r0;
  if (n <= 0) return;
  stage_weights<F>(ws, W, V, bias, K);
  __syncthreads();
  const int lr = threadIdx.x;
  const bool active = lr < n;
  const int r = r0 + lr;
  float dxr[F];
  if (active) {
    float p[F], xrow[F], go[F], dp[F];
#pragma unroll
    for (int q = 0; q < F / 4; ++q) {
      const float4 a = ldg4(prop + (size_t)r * F + 4 * q), b = ldg4(xr + (size_t)r * F + 4 * q),
                   c = ldg4(dout + (size_t)r * F + 4 * q);
      p[4 * q] = a.x, p[4 * q + 1] = a.y, p[4 * q + 2] = a.z, p[4 * q + 3] = a.w;
      xrow[4 * q] = b.x, xrow[4 * q + 1] = b.y, xrow[4 * q + 2] = b.z, xrow[4 * q + 3] = b.w;
      go[4 * q] = c.x, go[4 * q + 1] = c.y, go[4 * q + 2] = c.z, go[4 * q + 3] = c.w;
    }
    const float* cg = coeff + (int64_t)row_graph[r] * 2 * K;
    const float inv = 1.0f / (float)K;
#pragma unroll
    for (int j = 0; j < F; ++j) dp[j] = dxr[j] = 0.0f;
    for (int k = 0; k < K; ++k) {
      float u[F], v[F], dzk[F], tmp[F];
#pragma unroll
      for (int j = 0; j < F; ++j) u[j] = v[j] = 0.0f;
      apply_w<F>(u, p, Ws + k * F * F);
      apply_w<F>(v, xrow, Vs + k * F * F);
      const float a = __ldg(cg + k), b = __ldg(cg + K + k);
      float sa = 0.0f, sb = 0.0f;
#pragma unroll
      for (int j = 0; j < F; ++j) {
        const float z = fmaf(a, u[j], fmaf(b, v[j], bs[k * F + j]));
        dzk[j] = z > 0.0f ? go[j] * inv : 0.0f;
        sa = fmaf(u[j], dzk[j], sa);
        sb = fmaf(v[j], dzk[j], sb);
      }
      part[lr * 2 * K + k] = sa;
      part[lr * 2 * K + K + k] = sb;
      float* o = dz + (size_t)r * K * F + (size_t)k * F;
      float* oa = dza + (size_t)r * K * F + (size_t)k * F;
      float* ob = dzb + (size_t)r * K * F + (size_t)k * F;
#pragma unroll
      for (int q = 0; q < F / 4; ++q) {
        st4(o + 4 * q, make_float4(dzk[4 * q], dzk[4 * q + 1], dzk[4 * q + 2], dzk[4 * q + 3]));
        st4(oa + 4 * q, make_float4(a * dzk[4 * q], a * dzk[4 * q + 1], a * dzk[4 * q + 2], a * dzk[4 * q + 3]));
        st4(ob + 4 * q, make_float4(b * dzk[4 * q], b * dzk[4 * q + 1], b * dzk[4 * q + 2], b * dzk[4 * q + 3]));
      }
      apply_w_t<F>(tmp, dzk, Ws + k * F * F);
#pragma unroll
      for (int j = 0; j < F; ++j) dp[j] = fmaf(a, tmp[j], dp[j]);
      apply_w_t<F>(tmp, dzk, Vs + k * F * F);
#pragma unroll
      for (int j = 0; j < F; ++j) dxr[j] = fmaf(b, tmp[j], dxr[j]);
    }
    store_row<F>(buf0 + lr * LD, dp);
  }
  __syncthreads();
  {  // per-graph coefficient gradients, rows summed in order
    const int g0 = row_graph[r0], g1 = row_graph[r1 - 1];
    const int items = (g1 - g0 + 1) * 2 * K;
    for (int idx = threadIdx.x; idx < items; idx += blockDim.x) {
      const int gi = idx / (2 * K), c = idx - gi * 2 * K;
      const int a = graph_ptr[g0 + gi] - r0, b = graph_ptr[g0 + gi + 1] - r0;
      float s = 0.0f;
      for (int row = a; row < b; ++row) s += part[row * 2 * K + c];
      dcoeff[(int64_t)(g0 + gi) * 2 * K + c] = s;
    }
  }
  if (active) {
    float t[F];
    gather_row<F>(t, buf0, r0, colidx_t, vals_t, rowptr_t[r], rowptr_t[r + 1]);   // A_hat^T d(A_hat x)
    if (dx_root != nullptr) {
#pragma unroll
      for (int q = 0; q < F / 4; ++q) {
        st4(dx + (size_t)r * F + 4 * q, make_float4(t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]));
        st4(dx_root + (size_t)r * F + 4 * q, make_float4(dxr[4 * q], dxr[4 * q + 1], dxr[4 * q + 2], dxr[4 * q + 3]));
      }
    } else {
#pragma unroll
      for (int q = 0; q < F / 4; ++q)
        st4(dx + (size_t)r * F + 4 * q, make_float4(t[4 * q] + dxr[4 * q], t[4 * q + 1] + dxr[4 * q + 1],
                                                    t[4 * q + 2] + dxr[4 * q + 2], t[4 * q + 3] + dxr[4 * q + 3]));
    }
  }
}

// chunk quantum C / capacity as in fused_config (graph_tile.cuh); one thread per row, at most 512 (forward)
// or 256 (backward: ~150 live registers per row at F = 16) threads
static FusedCfg arma_config(int F, int K, int max_nodes, bool bwd) {
  FusedCfg c{false, 0, 0, 0, 0};
  if (!(F == 4 || F == 8 || F == 16)) return c;
  if (max_nodes < 1) max_nodes = 1;
  const int maxT = bwd ? 256 : 512;
  int C = (max_nodes + 31) / 32 * 32;
  if (C < 64) C = 64;
  if (C > 256) C = 256;
  while (C > 32 && C + max_nodes - 1 > maxT) C -= 32;
  const int cap = (C + max_nodes - 1 + 31) / 32 * 32;
  if (cap > maxT) return c;
  size_t smem = (size_t)(bwd ? 1 : 2) * cap * (F + 4) * sizeof(float);
  if (bwd) smem += (size_t)cap * 2 * K * sizeof(float);
  smem += (size_t)(2 * K * F * F + K * F) * sizeof(float);   // staged W | V | bias
  if (smem > 220 * 1024) return c;
  c.ok = true, c.C = C, c.cap = cap, c.threads = cap, c.smem = smem;
  return c;
}

#define FETA_ARMA_DISPATCH(F_, ...)                         \
  switch (F_) {                                             \
    case 4: { constexpr int FF = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int FF = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int FF = 16; __VA_ARGS__; } break; \
    default: break;                                         \
  }

}  // namespace feta

using namespace feta;

static int arma_check(int64_t R, int64_t G, int K, int F, const void* a, const void* b, const void* c, const void* d) {
  FETA_REQUIRE(R >= 0 && G >= 1 && K >= 1 && K <= 64, "arma: bad sizes R=%lld G=%lld K=%d", (long long)R,
               (long long)G, K);
  FETA_REQUIRE(F == 4 || F == 8 || F == 16, "arma: in_channels == out_channels must be 4, 8 or 16 (got %d)", F);
  FETA_REQUIRE(a && b && c && d, "arma: NULL pointer argument");
  return FETA_OK;
}

extern "C" int feta_arma_fwd(const float* x, const float* x_root, const int32_t* rowptr, const int32_t* colidx,
                             const float* vals, const int32_t* graph_ptr, const int32_t* row_graph,
                             int32_t* plan_meta, const float* coeff, const float* init_weight,
                             const float* root_weight, const float* bias, float* out, float* prop, int64_t R,
                             int64_t G, int K, int F, int max_nodes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (R == 0) return FETA_OK;
  int rc = arma_check(R, G, K, F, x, rowptr, coeff, out);
  if (rc) return rc;
  FETA_REQUIRE(graph_ptr && row_graph && init_weight && root_weight, "arma_fwd: NULL pointer argument");
  if (x_root == nullptr) x_root = x;
  const FusedCfg cfg = arma_config(F, K, max_nodes, false);
  if (!cfg.ok) {
    set_last_error("arma_fwd: largest graph (%d rows) does not fit one CTA", max_nodes);
    return FETA_EUNSUPPORTED;
  }
  const unsigned grid = (unsigned)ceil_div(R, cfg.C);
  FETA_ARMA_DISPATCH(F, {
    FETA_CUDA(cudaFuncSetAttribute(arma_fwd_kernel<FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    arma_fwd_kernel<FF><<<grid, cfg.threads, cfg.smem, st>>>(x, x_root, rowptr, colidx, vals, graph_ptr, row_graph,
                                                             coeff, init_weight, root_weight, bias, out, prop, R, K,
                                                             cfg.C, cfg.cap, plan_meta, G, max_nodes);
  });
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_arma_bwd(const float* dout, const float* prop, const float* x_root, const int32_t* rowptr_t,
                             const int32_t* colidx_t, const float* vals_t, const int32_t* graph_ptr,
                             const int32_t* row_graph, int32_t* plan_meta, const float* coeff,
                             const float* init_weight, const float* root_weight, const float* bias, float* dx,
                             float* dx_root, float* dz, float* dza, float* dzb, float* dcoeff, int64_t R, int64_t G,
                             int K, int F, int max_nodes, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  if (R == 0) return FETA_OK;
  int rc = arma_check(R, G, K, F, dout, prop, coeff, dx);
  if (rc) return rc;
  FETA_REQUIRE(x_root && rowptr_t && graph_ptr && row_graph && init_weight && root_weight && dz && dza && dzb && dcoeff,
               "arma_bwd: NULL pointer argument");
  const FusedCfg cfg = arma_config(F, K, max_nodes, true);
  if (!cfg.ok) {
    set_last_error("arma_bwd: largest graph (%d rows) does not fit one CTA", max_nodes);
    return FETA_EUNSUPPORTED;
  }
  const unsigned grid = (unsigned)ceil_div(R, cfg.C);
  FETA_ARMA_DISPATCH(F, {
    FETA_CUDA(cudaFuncSetAttribute(arma_bwd_kernel<FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem));
    arma_bwd_kernel<FF><<<grid, cfg.threads, cfg.smem, st>>>(dout, prop, x_root, rowptr_t, colidx_t, vals_t, graph_ptr,
                                                             row_graph, coeff, init_weight, root_weight, bias, dx,
                                                             dx_root, dz, dza, dzb, dcoeff, R, K, cfg.C, cfg.cap,
                                                             plan_meta, G, max_nodes);
  });
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}
