// A7: GPU batch builder (see include/feta_b200.h).
//
// Replaces the per-graph / per-node Python loops of GraphDataset_{v2,sbm,ogb}.collate_fn
// (transformer/data.py:161-225, :277-344, :394-460) for a dataset that has been packed onto the
// device once: node features, local edge lists, dense per-graph PE blocks and degree vectors
// concatenated with prefix-sum pointers.  Pure integer/index work + copies; outputs are
// bit-identical to the reference's host collate.
#include "common.cuh"

namespace feta {

constexpr int kColThreads = 256;

static inline unsigned col_grid(int64_t n) {
  int64_t b = ceil_div(n > 0 ? n : 1, kColThreads);
  const int64_t cap = (int64_t)kNumSMs * 32;
  return (unsigned)(b < cap ? b : cap);
}

// largest b with ptr[b] <= i   (ptr is [B+1], non-decreasing, ptr[B] > i)
__device__ __forceinline__ int owner_of(const int64_t* __restrict__ ptr, int B, int64_t i) {
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (ptr[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kColThreads) collate_nodes_kernel(const int64_t* __restrict__ out_node_ptr,
                                                                   int64_t* __restrict__ batch_indices,
                                                                   int64_t* __restrict__ feature_indices, int B,
                                                                   int64_t N) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = owner_of(out_node_ptr, B, i);
    batch_indices[i] = b;                        // data.py:220
    feature_indices[2 * i] = b;                  // data.py:218
    feature_indices[2 * i + 1] = i - out_node_ptr[b];
  }
}

__global__ void __launch_bounds__(kColThreads) collate_mask_kernel(const int64_t* __restrict__ out_node_ptr,
                                                                  uint8_t* __restrict__ mask, int B, int nmax) {
  const int64_t total = (int64_t)B * nmax;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / nmax), n = (int)(idx - (int64_t)b * nmax);
    mask[idx] = n >= (out_node_ptr[b + 1] - out_node_ptr[b]);  // data.py:210  mask[i, g_len:] = True
  }
}

__global__ void __launch_bounds__(kColThreads) collate_edges_kernel(
    const int64_t* __restrict__ graph_ids, const int64_t* __restrict__ ds_edge_ptr,
    const int64_t* __restrict__ ds_edge_index, int64_t ds_E, const int64_t* __restrict__ out_node_ptr,
    const int64_t* __restrict__ out_edge_ptr, int64_t* __restrict__ edge_indices, int B, int64_t E) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int b = owner_of(out_edge_ptr, B, e);
    const int64_t src = ds_edge_ptr[graph_ids[b]] + (e - out_edge_ptr[b]);
    const int64_t off = out_node_ptr[b];         // data.py:219  g.edge_index + node_offset
    edge_indices[e] = ds_edge_index[src] + off;
    edge_indices[E + e] = ds_edge_index[ds_E + src] + off;
  }
}

// static-shape variant: edge_indices is [2, e_cap]; columns >= E hold the (-1, -1) padding the plan builder ignores
__global__ void __launch_bounds__(kColThreads) collate_edges_static_kernel(
    const int64_t* __restrict__ graph_ids, const int64_t* __restrict__ ds_edge_ptr,
    const int64_t* __restrict__ ds_edge_index, int64_t ds_E, const int64_t* __restrict__ out_node_ptr,
    const int64_t* __restrict__ out_edge_ptr, int64_t* __restrict__ edge_indices, int B, int64_t E, int64_t e_cap) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < e_cap; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t s = -1, t = -1;
    if (e < E) {
      const int b = owner_of(out_edge_ptr, B, e);
      const int64_t src = ds_edge_ptr[graph_ids[b]] + (e - out_edge_ptr[b]);
      const int64_t off = out_node_ptr[b];
      s = ds_edge_index[src] + off;
      t = ds_edge_index[ds_E + src] + off;
    }
    edge_indices[e] = s;
    edge_indices[e_cap + e] = t;
  }
}

__global__ void __launch_bounds__(kColThreads) collate_pad_rows_kernel(const int64_t* __restrict__ graph_ids,
                                                                      const int64_t* __restrict__ ds_node_ptr,
                                                                      const float* __restrict__ src,
                                                                      float* __restrict__ dst, int B, int nmax, int C) {
  const int64_t total = (int64_t)B * nmax * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C);
    const int64_t bn = idx / C;
    const int b = (int)(bn / nmax), n = (int)(bn - (int64_t)b * nmax);
    const int64_t gid = graph_ids[b];
    const int64_t lo = ds_node_ptr[gid], len = ds_node_ptr[gid + 1] - lo;
    dst[idx] = n < len ? src[(lo + n) * C + c] : 0.0f;
  }
}

__global__ void __launch_bounds__(kColThreads) collate_pad_pe_kernel(const int64_t* __restrict__ graph_ids,
                                                                    const int64_t* __restrict__ ds_node_ptr,
                                                                    const int64_t* __restrict__ ds_pe_ptr,
                                                                    const float* __restrict__ pe_src,
                                                                    float* __restrict__ pe_dst, int B, int nmax) {
  const int64_t total = (int64_t)B * nmax * nmax;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % nmax);
    const int64_t bi = idx / nmax;
    const int b = (int)(bi / nmax), i = (int)(bi - (int64_t)b * nmax);
    const int64_t gid = graph_ids[b];
    const int64_t len = ds_node_ptr[gid + 1] - ds_node_ptr[gid];
    pe_dst[idx] = (i < len && j < len) ? pe_src[ds_pe_ptr[gid] + (int64_t)i * len + j] : 0.0f;
  }
}

}  // namespace feta

using namespace feta;

extern "C" int feta_collate_indices(const int64_t* graph_ids, const int64_t* ds_node_ptr, const int64_t* ds_edge_ptr,
                                    const int64_t* ds_edge_index, int64_t ds_E, const int64_t* out_node_ptr,
                                    const int64_t* out_edge_ptr, uint8_t* mask, int64_t* edge_indices,
                                    int64_t* batch_indices, int64_t* feature_indices, int B, int nmax, int64_t N,
                                    int64_t E, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  (void)ds_node_ptr;
  FETA_REQUIRE(B >= 1 && nmax >= 1 && N >= 0 && E >= 0, "collate_indices: bad sizes");
  FETA_REQUIRE(graph_ids && ds_edge_ptr && out_node_ptr && out_edge_ptr && mask, "collate_indices: NULL pointer");
  collate_mask_kernel<<<col_grid((int64_t)B * nmax), kColThreads, 0, st>>>(out_node_ptr, mask, B, nmax);
  FETA_LAUNCH_CHECK();
  if (N > 0) {
    FETA_REQUIRE(batch_indices && feature_indices, "collate_indices: NULL node outputs");
    collate_nodes_kernel<<<col_grid(N), kColThreads, 0, st>>>(out_node_ptr, batch_indices, feature_indices, B, N);
    FETA_LAUNCH_CHECK();
  }
  if (E > 0) {
    FETA_REQUIRE(edge_indices && ds_edge_index, "collate_indices: NULL edge pointers");
    collate_edges_kernel<<<col_grid(E), kColThreads, 0, st>>>(graph_ids, ds_edge_ptr, ds_edge_index, ds_E,
                                                              out_node_ptr, out_edge_ptr, edge_indices, B, E);
    FETA_LAUNCH_CHECK();
  }
  return FETA_OK;
}

extern "C" int feta_collate_edges_static(const int64_t* graph_ids, const int64_t* ds_edge_ptr,
                                         const int64_t* ds_edge_index, int64_t ds_E, const int64_t* out_node_ptr,
                                         const int64_t* out_edge_ptr, int64_t* edge_indices, int B, int64_t E,
                                         int64_t e_cap, void* stream_) {
  FETA_REQUIRE(B >= 1 && E >= 0 && e_cap >= E && e_cap >= 1, "collate_edges_static: bad sizes (E must fit e_cap)");
  FETA_REQUIRE(graph_ids && ds_edge_ptr && ds_edge_index && out_node_ptr && out_edge_ptr && edge_indices,
               "collate_edges_static: NULL pointer");
  collate_edges_static_kernel<<<col_grid(e_cap), kColThreads, 0, (cudaStream_t)stream_>>>(
      graph_ids, ds_edge_ptr, ds_edge_index, ds_E, out_node_ptr, out_edge_ptr, edge_indices, B, E, e_cap);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_collate_pad_rows(const int64_t* graph_ids, const int64_t* ds_node_ptr, const float* src, float* dst,
                                     int B, int nmax, int C, void* stream_) {
  FETA_REQUIRE(graph_ids && ds_node_ptr && src && dst && B >= 1 && nmax >= 1 && C >= 1, "collate_pad_rows: bad argument");
  collate_pad_rows_kernel<<<col_grid((int64_t)B * nmax * C), kColThreads, 0, (cudaStream_t)stream_>>>(
      graph_ids, ds_node_ptr, src, dst, B, nmax, C);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_collate_pad_pe(const int64_t* graph_ids, const int64_t* ds_node_ptr, const int64_t* ds_pe_ptr,
                                   const float* pe_src, float* pe_dst, int B, int nmax, void* stream_) {
  FETA_REQUIRE(graph_ids && ds_node_ptr && ds_pe_ptr && pe_src && pe_dst && B >= 1 && nmax >= 1,
               "collate_pad_pe: bad argument");
  collate_pad_pe_kernel<<<col_grid((int64_t)B * nmax * nmax), kColThreads, 0, (cudaStream_t)stream_>>>(
      graph_ids, ds_node_ptr, ds_pe_ptr, pe_src, pe_dst, B, nmax);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Static-shape batches (engine.GraphedTrainStep): everything DiffTransformerEncoderGenGCN.forward_static derives from
// the padding mask and the packed edge list -- graph sizes, their prefix, the packed-id -> padded-slot map of every
// edge endpoint, the segment bounds of the coefficient pooling and the real-row weights -- in two launches instead
// of ~15 tiny tensor ops (each a ~2 us link of the step's dependent chain).
// ---------------------------------------------------------------------------------------------------------------
namespace feta {

// one CTA: lens[b] = # un-masked positions, node_end = inclusive prefix; seg bounds; real-row weights
__global__ void __launch_bounds__(1024) static_sizes_kernel(const uint8_t* __restrict__ mask, int B, int nmax, int H,
                                                            int32_t* __restrict__ node_end, int32_t* __restrict__ lens,
                                                            int32_t* __restrict__ slot_ptr, int32_t* __restrict__ seg_lo,
                                                            int32_t* __restrict__ seg_hi, float* __restrict__ real) {
  extern __shared__ int32_t sl[];                       // [B] lens, then their inclusive prefix
  __shared__ int32_t wsum[32];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, nw = blockDim.x >> 5;
  for (int b = warp; b < B; b += nw) {                  // warp per graph
    int c = 0;
    for (int j = lane; j < nmax; j += 32) c += mask[(size_t)b * nmax + j] == 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) sl[b] = c, lens[b] = c;
  }
  __syncthreads();
  // inclusive scan of sl[0..B) in chunks of blockDim.x
  int32_t carry = 0;
  for (int b0 = 0; b0 < B; b0 += blockDim.x) {
    const int b = b0 + t;
    int32_t v = b < B ? sl[b] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) wsum[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int32_t w = lane < nw ? wsum[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int32_t u = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += u;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const int32_t incl = v + (warp > 0 ? wsum[warp - 1] : 0) + carry;
    if (b < B) node_end[b] = incl;
    carry += wsum[nw - 1];
    __syncthreads();
  }
  for (int b = t; b < B; b += blockDim.x) slot_ptr[b] = b * nmax;
  for (int g = t; g < H * B; g += blockDim.x) {
    seg_lo[g] = g * nmax;
    seg_hi[g] = g * nmax + sl[g % B];
  }
  for (int idx = t; idx < nmax * B; idx += blockDim.x) {          // real [nmax, B]
    const int i = idx / B, b = idx - i * B;
    real[idx] = mask[(size_t)b * nmax + i] == 0 ? 1.0f : 0.0f;
  }
}

// packed node id -> padded slot id (b * nmax + i) of every edge endpoint; (-1, -1) padding columns stay -1
template <typename IT>
__global__ void __launch_bounds__(256) static_edges_kernel(const IT* __restrict__ ei, int64_t ecap,
                                                           const int32_t* __restrict__ node_end,
                                                           const int32_t* __restrict__ lens, int B, int nmax, int heads,
                                                           int64_t* __restrict__ out) {
  const int64_t n = 2 * ecap;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = (int64_t)ei[idx];
    int64_t r = -1;
    if (v >= 0) {
      int lo = 0, hi = B;                               // first b with node_end[b] > v  (searchsorted right=True)
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(node_end + mid) > v) hi = mid;
        else lo = mid + 1;
      }
      const int b = lo < B - 1 ? lo : B - 1;
      r = v - ((int64_t)__ldg(node_end + b) - __ldg(lens + b)) + (int64_t)b * nmax;
    }
    const int64_t row = idx / ecap, col = idx - row * ecap;
    for (int h = 0; h < heads; ++h)                     // per-head tiling (opt-in): head h's copy sits h * B * nmax further
      out[row * ecap * heads + (int64_t)h * ecap + col] = r < 0 ? -1 : r + (int64_t)h * B * nmax;
  }
}

}  // namespace feta

extern "C" int feta_static_context(const uint8_t* mask, const void* edge_index, int edge_dtype, int64_t ecap, int B,
                                   int nmax, int H, int tile_heads, int64_t* ei_out, int32_t* node_end, int32_t* lens,
                                   int32_t* slot_ptr, int32_t* seg_lo, int32_t* seg_hi, float* real, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(B >= 1 && nmax >= 1 && H >= 1 && ecap >= 0, "static_context: bad sizes");
  FETA_REQUIRE(mask && ei_out && node_end && lens && slot_ptr && seg_lo && seg_hi && real && (edge_index || ecap == 0),
               "static_context: NULL pointer argument");
  FETA_REQUIRE(edge_dtype == FETA_DT_I32 || edge_dtype == FETA_DT_I64, "static_context: edge dtype must be int32/int64");
  FETA_REQUIRE((size_t)B * sizeof(int32_t) <= 48 * 1024, "static_context: B=%d too large", B);
  feta::static_sizes_kernel<<<1, 1024, (size_t)B * sizeof(int32_t), st>>>(mask, B, nmax, H, node_end, lens, slot_ptr,
                                                                        seg_lo, seg_hi, real);
  FETA_LAUNCH_CHECK();
  if (ecap > 0) {
    const int heads = tile_heads ? H : 1;
    int64_t blocks = feta::ceil_div(2 * ecap, 256);
    if (blocks > 4 * feta::kNumSMs) blocks = 4 * feta::kNumSMs;
    if (edge_dtype == FETA_DT_I32)
      feta::static_edges_kernel<int32_t><<<(unsigned)blocks, 256, 0, st>>>((const int32_t*)edge_index, ecap, node_end,
                                                                          lens, B, nmax, heads, ei_out);
    else
      feta::static_edges_kernel<int64_t><<<(unsigned)blocks, 256, 0, st>>>((const int64_t*)edge_index, ecap, node_end,
                                                                          lens, B, nmax, heads, ei_out);
    FETA_LAUNCH_CHECK();
  }
  return FETA_OK;
}
