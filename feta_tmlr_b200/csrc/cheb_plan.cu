// A2: CSR build of the scaled Laplacian + graph segmentation (see include/feta_b200.h).
//
// Replaces ChebConvDynamic.__norm__ (transformer/ChebNetDynamic.py:108-130) and the
// unique/return_counts segmenting of :148.  Everything is integer/index work except the
// edge weight -(2/lambda_max) * deg[s]^-1/2 * deg[t]^-1/2.  Deterministic: entries of a CSR
// row are emitted in input edge order (the atomically claimed slots are re-ranked by edge id).
#include "common.cuh"

namespace feta {

constexpr int kThreads = 256;

__global__ void plan_init_meta_kernel(int32_t* meta) {
  if (threadIdx.x < FETA_META_WORDS) {
    int32_t v = 0;
    if (threadIdx.x == FETA_META_SORTED || threadIdx.x == FETA_META_BLOCKDIAG) v = 1;
    meta[threadIdx.x] = v;
  }
}

__global__ void __launch_bounds__(kThreads) plan_count_kernel(const int64_t* __restrict__ ei, int64_t E,
                                                             int64_t R, int32_t* outdeg, int32_t* indeg,
                                                             int32_t* meta, bool keep_self) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = ei[e], t = ei[E + e];
    if (s == -1 && t == -1) continue;  // padding column of a fixed-width edge list
    if (s < 0 || s >= R || t < 0 || t >= R) {
      meta[6] = 1;  // out-of-range endpoint: reported by the host wrapper
      continue;
    }
    if (s != t || keep_self) {  // remove_self_loops, ChebNetDynamic.py:113 (gcn_norm keeps them)
      atomicAdd(&outdeg[s], 1);
      atomicAdd(&indeg[t], 1);
    }
  }
}

__global__ void __launch_bounds__(kThreads) plan_fill_kernel(const int64_t* __restrict__ ei, int64_t E,
                                                            int64_t R, const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ rowptr_t,
                                                            int32_t* cursor, int32_t* cursor_t,
                                                            int32_t* tmp, int32_t* tmp_t, bool keep_self) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = ei[e], t = ei[E + e];
    if (s < 0 || s >= R || t < 0 || t >= R || (s == t && !keep_self)) continue;
    tmp[rowptr[t] + atomicAdd(&cursor[t], 1)] = (int32_t)e;
    tmp_t[rowptr_t[s] + atomicAdd(&cursor_t[s], 1)] = (int32_t)e;
  }
}

// One thread per edge: its final slot inside the row is the number of row-mates with a smaller
// edge id, which restores input order whatever order the atomics claimed the slots in.
__global__ void __launch_bounds__(kThreads) plan_rank_kernel(
    const int64_t* __restrict__ ei, int64_t E, int64_t R, float two_over_lambda, int norm_mode,
    const int32_t* __restrict__ outdeg, const int32_t* __restrict__ indeg, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ rowptr_t, const int32_t* __restrict__ tmp,
    const int32_t* __restrict__ tmp_t, int32_t* __restrict__ colidx, float* __restrict__ vals,
    int32_t* __restrict__ colidx_t, float* __restrict__ vals_t, int32_t* meta) {
  int32_t maxdeg = 0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E;
       e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = ei[e], t = ei[E + e];
    const bool gcn = norm_mode == FETA_NORM_GCN;
    if (s < 0 || s >= R || t < 0 || t >= R || (s == t && !gcn)) continue;
    // get_laplacian('sym'): deg over the SOURCE index, deg^-1/2 with inf -> 0
    // gcn_norm (ARMAConvDynamic, ChebNetDynamic.py:301-305): deg over the TARGET index, self-loops kept
    const int32_t ds = gcn ? indeg[s] : outdeg[s], dt = gcn ? indeg[t] : outdeg[t];
    const float is = ds > 0 ? 1.0f / sqrtf((float)ds) : 0.0f;
    const float it = dt > 0 ? 1.0f / sqrtf((float)dt) : 0.0f;
    const float w = gcn ? is * it : -(is * it) * two_over_lambda;  // (2 * -w) / lambda_max, :122
    {
      const int32_t a = rowptr[t], b = rowptr[t + 1];
      int32_t rank = 0;
      for (int32_t p = a; p < b; ++p) rank += (tmp[p] < (int32_t)e);
      colidx[a + rank] = (int32_t)s;
      vals[a + rank] = w;
      maxdeg = max(maxdeg, b - a);
    }
    {
      const int32_t a = rowptr_t[s], b = rowptr_t[s + 1];
      int32_t rank = 0;
      for (int32_t p = a; p < b; ++p) rank += (tmp_t[p] < (int32_t)e);
      colidx_t[a + rank] = (int32_t)t;
      vals_t[a + rank] = w;
      maxdeg = max(maxdeg, b - a);
    }
  }
  if (maxdeg > 0) atomicMax(&meta[FETA_META_MAX_DEG], maxdeg);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) plan_batch_flags_kernel(const T* __restrict__ batch, int64_t R,
                                                                   int32_t* flags, int32_t* meta) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R;
       i += (int64_t)gridDim.x * blockDim.x) {
    int32_t f = 0;
    if (i > 0) {
      const T a = batch[i - 1], b = batch[i];
      f = (a != b);
      if (b < a) meta[FETA_META_SORTED] = 0;
    }
    flags[i] = f;
  }
}

// seg_excl = exclusive scan of flags; run index of row i = seg_excl[i] + flags[i]
__global__ void __launch_bounds__(kThreads) plan_graph_ptr_kernel(const int32_t* __restrict__ flags,
                                                                 const int32_t* __restrict__ seg_excl,
                                                                 int64_t R, int64_t G, int32_t* graph_ptr,
                                                                 int32_t* row_graph, int32_t* meta) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t g = seg_excl[i] + flags[i];
    row_graph[i] = g;
    if ((i == 0 || flags[i]) && g <= G) graph_ptr[g] = (int32_t)i;
    if (i == R - 1) meta[FETA_META_NUM_GRAPHS] = g + 1;
  }
}

__global__ void __launch_bounds__(kThreads) plan_graph_tail_kernel(int64_t R, int64_t G, int32_t* graph_ptr,
                                                                  const int32_t* meta) {
  // graph_ptr[num_found .. G] = R  (also the closing sentinel)
  const int32_t found = R > 0 ? meta[FETA_META_NUM_GRAPHS] : 0;
  for (int64_t g = found + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= G;
       g += (int64_t)gridDim.x * blockDim.x)
    graph_ptr[g] = (int32_t)R;
}

__global__ void __launch_bounds__(kThreads) plan_single_graph_kernel(int64_t R, int32_t* graph_ptr,
                                                                    int32_t* row_graph, int32_t* meta) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R;
       i += (int64_t)gridDim.x * blockDim.x)
    row_graph[i] = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    graph_ptr[0] = 0;
    graph_ptr[1] = (int32_t)R;
    meta[FETA_META_NUM_GRAPHS] = 1;
  }
}

__global__ void __launch_bounds__(kThreads) plan_check_kernel(const int32_t* __restrict__ rowptr,
                                                             const int32_t* __restrict__ colidx,
                                                             const int32_t* __restrict__ graph_ptr,
                                                             const int32_t* __restrict__ row_graph,
                                                             int64_t R, int64_t G, int32_t* meta) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < R;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int32_t g = row_graph[r];
    if (g >= G) continue;  // more runs than capacity: host raises on NUM_GRAPHS mismatch
    const int32_t lo = graph_ptr[g], hi = graph_ptr[g + 1];
    if (r == lo) atomicMax(&meta[FETA_META_MAX_NODES], hi - lo);
    bool ok = true;
    for (int32_t p = rowptr[r]; p < rowptr[r + 1]; ++p) {
      const int32_t c = colidx[p];
      ok = ok && (c >= lo) && (c < hi);
    }
    if (!ok) meta[FETA_META_BLOCKDIAG] = 0;
  }
}

static inline unsigned grid_for(int64_t n) {
  int64_t b = ceil_div(n > 0 ? n : 1, kThreads);
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (unsigned)(b < cap ? b : cap);
}

}  // namespace feta

using namespace feta;

extern "C" size_t feta_cheb_plan_workspace_bytes(int64_t R, int64_t E) {
  if (R < 0) R = 0;
  if (E < 0) E = 0;
  size_t b = 0;
  b += 4 * align_up((size_t)(R + 1) * 4, 256);                                 // outdeg, indeg, cursor, cursor_t
  b += 2 * align_up((size_t)(E + 1) * 4, 256);                                 // tmp, tmp_t
  b += 2 * align_up((size_t)(R + 1) * 4, 256);                                 // flags, seg
  b += align_up(scan_scratch_ints(R + 1) * 4, 256);
  return b + 1024;
}

extern "C" int feta_cheb_plan_build(const int64_t* edge_index, int64_t E, const void* batch, int batch_dtype,
                                    int64_t R, int64_t G, float lambda_max, int32_t* rowptr, int32_t* colidx,
                                    float* vals, int32_t* rowptr_t, int32_t* colidx_t, float* vals_t,
                                    int32_t* graph_ptr, int32_t* row_graph, int32_t* meta, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
  return feta_graph_plan_build(edge_index, E, batch, batch_dtype, R, G, FETA_NORM_CHEB_SYM, lambda_max, rowptr,
                               colidx, vals, rowptr_t, colidx_t, vals_t, graph_ptr, row_graph, meta, workspace,
                               workspace_bytes, stream_);
}

extern "C" int feta_graph_plan_build(const int64_t* edge_index, int64_t E, const void* batch, int batch_dtype,
                                     int64_t R, int64_t G, int norm_mode, float lambda_max, int32_t* rowptr,
                                     int32_t* colidx, float* vals, int32_t* rowptr_t, int32_t* colidx_t,
                                     float* vals_t, int32_t* graph_ptr, int32_t* row_graph, int32_t* meta,
                                     void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FETA_REQUIRE(norm_mode == FETA_NORM_CHEB_SYM || norm_mode == FETA_NORM_GCN, "plan_build: bad norm_mode %d",
               norm_mode);
  const bool keep_self = norm_mode == FETA_NORM_GCN;
  FETA_REQUIRE(R >= 0 && E >= 0 && G >= 1, "plan_build: bad sizes R=%lld E=%lld G=%lld", (long long)R,
               (long long)E, (long long)G);
  FETA_REQUIRE(R < (1ll << 31) && E < (1ll << 31), "plan_build: R/E exceed int32 index range");
  FETA_REQUIRE(E == 0 || edge_index, "plan_build: edge_index is NULL");
  FETA_REQUIRE(rowptr && rowptr_t && graph_ptr && meta && (R == 0 || row_graph),
               "plan_build: NULL output pointer");
  FETA_REQUIRE(E == 0 || (colidx && vals && colidx_t && vals_t), "plan_build: NULL CSR output pointer");
  FETA_REQUIRE(lambda_max > 0.0f, "plan_build: lambda_max must be positive");
  FETA_REQUIRE(batch_dtype >= FETA_DT_I64 && batch_dtype <= FETA_DT_F64, "plan_build: bad batch dtype %d",
               batch_dtype);
  if (workspace_bytes < feta_cheb_plan_workspace_bytes(R, E) || !workspace) {
    set_last_error("plan_build: workspace too small (%zu < %zu)", workspace_bytes,
                   feta_cheb_plan_workspace_bytes(R, E));
    return FETA_EWORKSPACE;
  }
  Arena ar(workspace, workspace_bytes);
  int32_t* outdeg = ar.take<int32_t>(R + 1);
  int32_t* indeg = ar.take<int32_t>(R + 1);
  int32_t* cursor = ar.take<int32_t>(R + 1);
  int32_t* cursor_t = ar.take<int32_t>(R + 1);
  int32_t* tmp = ar.take<int32_t>(E + 1);
  int32_t* tmp_t = ar.take<int32_t>(E + 1);
  int32_t* flags = ar.take<int32_t>(R + 1);
  int32_t* seg = ar.take<int32_t>(R + 1);
  int32_t* scratch = ar.take<int32_t>(scan_scratch_ints(R + 1));
  FETA_REQUIRE(scratch != nullptr, "plan_build: workspace carve failed");

  plan_init_meta_kernel<<<1, 32, 0, stream>>>(meta);
  FETA_LAUNCH_CHECK();
  // outdeg, indeg, cursor, cursor_t are contiguous arena blocks
  FETA_CUDA(cudaMemsetAsync(outdeg, 0, (size_t)((char*)tmp - (char*)outdeg), stream));
  if (E > 0) {
    plan_count_kernel<<<grid_for(E), kThreads, 0, stream>>>(edge_index, E, R, outdeg, indeg, meta, keep_self);
    FETA_LAUNCH_CHECK();
  }
  int rc = exclusive_scan_i32(indeg, rowptr, R, 1, scratch, stream);
  if (rc) return rc;
  rc = exclusive_scan_i32(outdeg, rowptr_t, R, 1, scratch, stream);
  if (rc) return rc;
  if (E > 0) {
    plan_fill_kernel<<<grid_for(E), kThreads, 0, stream>>>(edge_index, E, R, rowptr, rowptr_t, cursor, cursor_t,
                                                          tmp, tmp_t, keep_self);
    FETA_LAUNCH_CHECK();
    plan_rank_kernel<<<grid_for(E), kThreads, 0, stream>>>(edge_index, E, R, 2.0f / lambda_max, norm_mode, outdeg,
                                                          indeg, rowptr, rowptr_t, tmp, tmp_t, colidx, vals,
                                                          colidx_t, vals_t, meta);
    FETA_LAUNCH_CHECK();
  }
  FETA_CUDA(cudaMemcpyAsync(meta + FETA_META_NNZ, rowptr + R, sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));

  if (batch == nullptr || R == 0) {
    FETA_REQUIRE(G >= 1, "plan_build: G must be >= 1");
    plan_single_graph_kernel<<<grid_for(R), kThreads, 0, stream>>>(R, graph_ptr, row_graph, meta);
    FETA_LAUNCH_CHECK();
    if (G > 1 || R == 0) {
      plan_graph_tail_kernel<<<grid_for(G + 1), kThreads, 0, stream>>>(R, G, graph_ptr, meta);
      FETA_LAUNCH_CHECK();
    }
  } else {
    switch (batch_dtype) {
      case FETA_DT_I64:
        plan_batch_flags_kernel<int64_t><<<grid_for(R), kThreads, 0, stream>>>((const int64_t*)batch, R, flags, meta);
        break;
      case FETA_DT_I32:
        plan_batch_flags_kernel<int32_t><<<grid_for(R), kThreads, 0, stream>>>((const int32_t*)batch, R, flags, meta);
        break;
      case FETA_DT_F32:
        plan_batch_flags_kernel<float><<<grid_for(R), kThreads, 0, stream>>>((const float*)batch, R, flags, meta);
        break;
      default:
        plan_batch_flags_kernel<double><<<grid_for(R), kThreads, 0, stream>>>((const double*)batch, R, flags, meta);
        break;
    }
    FETA_LAUNCH_CHECK();
    rc = exclusive_scan_i32(flags, seg, R, 0, scratch, stream);
    if (rc) return rc;
    plan_graph_ptr_kernel<<<grid_for(R), kThreads, 0, stream>>>(flags, seg, R, G, graph_ptr, row_graph, meta);
    FETA_LAUNCH_CHECK();
    plan_graph_tail_kernel<<<grid_for(G + 1), kThreads, 0, stream>>>(R, G, graph_ptr, meta);
    FETA_LAUNCH_CHECK();
  }
  if (R > 0) {
    plan_check_kernel<<<grid_for(R), kThreads, 0, stream>>>(rowptr, colidx, graph_ptr, row_graph, R, G, meta);
    FETA_LAUNCH_CHECK();
  }
  return FETA_OK;
}
