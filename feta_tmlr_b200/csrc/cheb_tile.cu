// A1/A3: ChebConvDynamic forward / backward as TILE kernels on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Design.  A tile is up to 128*MT consecutive rows made of at most J whole graphs
// (block-diagonal operator: a tile is closed under neighbours).  One compute thread owns one row; two control
// warps drive the asynchronous engines:
//   * loader warp: walks the graph list (persistent grid over row windows, greedy packing of whole graphs into
//     tiles), and brings everything a tile needs into shared memory with cp.async.bulk (TMA, SASS UBLKCP)
//     completing on mbarriers: the x slab (one copy), the Theta blocks of the tile's graphs (one copy when the
//     coefficient tensor is contiguous), the CSR slice (two copies) and the rowptr slice;
//   * MMA warp: one elected lane issues tcgen05.mma.kind::tf32 for the filter application
//         D[rows, J*F] (+)= T_k[rows, F] . [Theta_k(g_0) | ... | Theta_k(g_{J-1})]
//     -- every row is multiplied by the filters of ALL graphs of the tile (N-stacking; the tensor pipe has the
//     headroom) and the epilogue keeps the F columns of the row's own graph.  fp32-grade accuracy through the
//     3-term TF32 split (hi.hi + hi.lo + lo.hi), accumulators in TMEM across all K orders;
//   * compute threads: T_k = 2 L T_{k-1} - T_{k-2} by a float4 gather over the staged CSR slice out of the
//     UMMA-canonical T_{k-1} slab (8x16-byte core matrices: lane-per-row 16-byte accesses are conflict-free),
//     write T_k (raw fp32 = the "hi" operand, the tensor core reads its top 19 bits) and its remainder "lo",
//     then hand the slab to the MMA warp through a named barrier.  The FMAs of the filter application, the
//     operand splitting per MMA fragment and the Theta broadcasts of the warp-per-graph kernel (cheb_warp.cu,
//     1686 warp instructions per 23-row graph) leave the SM's issue slots.
// The backward kernels reuse the skeleton: dx = Clenshaw recursion over D_k = dOut . Theta_k^T (all K products
// issued at once into K TMEM column blocks), dTheta_k[g] = T_k^T dOut reduced per graph in shared memory.
#include <stdlib.h>

#include "common.cuh"
#include "graph_tile.cuh"
#include "umma.cuh"

namespace feta {
namespace tile {

using namespace tc;

constexpr int kEdgeCap = 384;   // staged CSR entries per 128 rows (larger slices are read through L2 instead)
constexpr int kRpPad = 8;

template <int F>
struct Geo {
  static constexpr int KC = F / 4;     // 16-byte K-cores per row
  static constexpr int J = 64 / F;     // graphs per tile  (N = J*F = 64 accumulator columns per M-tile)
  static constexpr int N = J * F;
};

// byte offset of (row, kcore q) in a canonical K-major slab with KC cores per row
template <int KC>
__device__ __forceinline__ uint32_t canon_row(int row) {
  return (uint32_t)(((row >> 3) * KC) << 7) + (uint32_t)((row & 7) << 4);
}

struct TileDesc {          // written by the loader warp, one slot per CSR stage
  int32_t r0, n, g0, cnt;  // first row, rows, first graph, graphs (cnt == 0: no more tiles)
  int32_t e_base;          // global index of staged colidx/vals element 0, or -1 when the slice is not staged
  int32_t rp_base;         // global index of staged rowptr element 0
  int32_t gp[10];          // graph_ptr[g0 .. g0+cnt]
};

struct Smem {              // byte offsets of the carve (all multiples of 128)
  uint32_t stage, stage_bytes, xd, xd2, th, csr, ahi, alo, bhi, blo, gbuf, misc, total;
};

// One landing stage = everything the loader brings in for a tile: [xd | xd2 | th | colidx | vals | rowptr].
template <int F, int MT>
__host__ __device__ inline Smem carve(int K, int n_hi, int n_lo, bool with_b, bool with_xd2, bool with_th,
                                      int gbuf_bytes) {
  constexpr int ROWS = 128 * MT;
  Smem s;
  uint32_t o = 0;
  s.xd = o;  o += ROWS * F * 4;
  s.xd2 = o; o += with_xd2 ? ROWS * F * 4 : 0;
  s.th = o;  o += with_th ? Geo<F>::J * K * F * F * 4 : 0;
  s.csr = o; o += (uint32_t)(kEdgeCap * MT * 8 + ((ROWS + kRpPad + 31) / 32 * 32) * 4);
  s.stage_bytes = o;
  s.stage = 0;
  o = 2 * s.stage_bytes;
  s.ahi = o; o += n_hi * ROWS * F * 4;
  s.alo = o; o += n_lo * ROWS * F * 4;
  s.bhi = o; o += with_b ? K * Geo<F>::N * F * 4 : 0;
  s.blo = o; o += with_b ? K * Geo<F>::N * F * 4 : 0;
  s.gbuf = o; o += gbuf_bytes;
  s.misc = o; o += 512;
  s.total = o;
  return s;
}

// misc block: mbarriers, tile descriptors, tmem slot, bias
struct Misc {
  uint64_t full[2], used[2], mma[4], a_ready[2];
  uint64_t pad_[6];
  TileDesc desc[2];
  uint32_t tmem;
  uint32_t pad2_[15];
  float bias[32];
};
static_assert(sizeof(TileDesc) == 64, "TileDesc is one 64-byte slot");
static_assert(sizeof(Misc) <= 512, "misc block");

__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) { mbar_wait_parity_sleep(smem_u32(bar), parity, 2000); }
__device__ __forceinline__ void wait_short(uint64_t* bar, uint32_t parity) { mbar_wait_parity_sleep(smem_u32(bar), parity, 100); }

// ------------------------------------------------------------------------------------------------
// loader warp: tile discovery + TMA.  `rows1` / `rows2` = the [R, F] operands staged per tile (x, dOut);
// the CSR is the one the recursion walks (target-grouped for forward / dTheta, source-grouped for dx).
// Two landing stages: the loads of tile t+1 are issued as soon as tile t-1 released its stage.
// ------------------------------------------------------------------------------------------------
template <int F, int MT>
__device__ __forceinline__ void loader_warp(unsigned char* smem, const Smem& L, Misc* M, const float* __restrict__ rows1,
                                            const float* __restrict__ rows2, const float* __restrict__ theta,
                                            int64_t sk, int64_t sg, const int32_t* __restrict__ rowptr,
                                            const int32_t* __restrict__ colidx, const float* __restrict__ vals,
                                            const int32_t* __restrict__ graph_ptr,
                                            const int32_t* __restrict__ row_graph, int64_t R, int64_t G, int K,
                                            int window) {
  constexpr int ROWS = 128 * MT, J = Geo<F>::J;
  const int lane = threadIdx.x & 31;
  const int nnz_total = __ldg(rowptr + R);
  const bool theta_contig = (sk == (int64_t)F * F) && (sg == (int64_t)K * F * F);
  const int64_t num_windows = (R + window - 1) / window;
  int t = 0;
  for (int64_t w = blockIdx.x; w < num_windows; w += gridDim.x) {
    const int64_t p0 = w * window, p1 = p0 + window;
    int gs, ge;
    {
      const int g = __ldg(row_graph + p0);
      gs = (__ldg(graph_ptr + g) == (int)p0) ? g : g + 1;
      if (p1 >= R) ge = (int)G;
      else {
        const int g1 = __ldg(row_graph + p1);
        ge = (__ldg(graph_ptr + g1) == (int)p1) ? g1 : g1 + 1;
      }
    }
    int g = gs;
    while (g < ge) {
      const int gi = g + lane;
      const int v = __ldg(graph_ptr + (gi <= (int)G ? gi : (int)G));
      const int r0 = __shfl_sync(0xffffffffu, v, 0);
      const bool fits = lane >= 1 && lane <= J && gi <= ge && (v - r0) <= ROWS;
      const unsigned m = __ballot_sync(0xffffffffu, fits) >> 1;      // bit i-1: graphs g .. g+i-1 fit
      int cnt = __ffs(~m) - 1;                                        // leading run of ones
      if (cnt <= 0) cnt = 1;   // a graph larger than the tile: excluded by the host + the plan guard
      const int r1 = __shfl_sync(0xffffffffu, v, cnt);
      const int n = r1 - r0;
      int e_lo = 0, e_hi = 0;
      if (lane == 0) e_lo = __ldg(rowptr + r0);
      if (lane == 1) e_hi = __ldg(rowptr + r1);
      e_lo = __shfl_sync(0xffffffffu, e_lo, 0);
      e_hi = __shfl_sync(0xffffffffu, e_hi, 1);
      const int a_lo = e_lo & ~3, a_hi = (e_hi + 3) & ~3;
      const bool staged = (a_hi - a_lo) <= kEdgeCap * MT && a_hi <= nnz_total && e_hi > e_lo;
      // rowptr slice [r0, r1]: bulk copies move multiples of 16 bytes; the last tile's tail (rowptr has exactly
      // R+1 entries) is copied by the lanes instead
      const int rp_lo = r0 & ~3;
      const int rp_hi = (r1 + 1 + 3) & ~3;
      const int rp_end = (int)((R + 1) & ~(int64_t)3);
      const int rp_tail = rp_hi > rp_end ? rp_end : rp_hi;       // bulk part: [rp_lo, rp_tail)
      const int s = t & 1;
      if (t >= 2) wait(&M->used[s], (uint32_t)(((t >> 1) - 1) & 1));   // the stage was released by tile t-2
      unsigned char* st = smem + L.stage + (size_t)s * L.stage_bytes;
      TileDesc* d = &M->desc[s];
      if (lane <= J) d->gp[lane] = v;
      if (rp_tail < rp_hi) {
        int32_t* rp_s = reinterpret_cast<int32_t*>(st + L.csr + kEdgeCap * MT * 8);
        for (int i = rp_tail + lane; i <= r1; i += 32) rp_s[i - rp_lo] = __ldg(rowptr + i);
      }
      if (lane == 0) {
        d->r0 = r0; d->n = n; d->g0 = g; d->cnt = cnt;
        d->e_base = staged ? a_lo : -1;
        d->rp_base = rp_lo;
      }
      __syncwarp();
      if (lane == 0) {
        const uint32_t row_bytes = (uint32_t)n * F * 4;
        const uint32_t th_bytes = theta ? (uint32_t)cnt * K * F * F * 4 : 0u;
        const uint32_t e_bytes = staged ? (uint32_t)(a_hi - a_lo) * 4u : 0u;
        const uint32_t rp_bytes = rp_tail > rp_lo ? (uint32_t)(rp_tail - rp_lo) * 4u : 0u;
        const uint32_t bar = smem_u32(&M->full[s]);
        mbar_arrive_expect_tx(bar, row_bytes * (rows2 ? 2u : 1u) + th_bytes + 2 * e_bytes + rp_bytes);
        bulk_g2s(smem_u32(st + L.xd), rows1 + (size_t)r0 * F, row_bytes, bar);
        if (rows2) bulk_g2s(smem_u32(st + L.xd2), rows2 + (size_t)r0 * F, row_bytes, bar);
        if (theta) {
          if (theta_contig) {
            bulk_g2s(smem_u32(st + L.th), theta + (int64_t)g * sg, th_bytes, bar);
          } else {
            for (int j = 0; j < cnt; ++j)
              for (int k = 0; k < K; ++k)
                bulk_g2s(smem_u32(st + L.th + (uint32_t)(j * K + k) * F * F * 4),
                         theta + (int64_t)(g + j) * sg + (int64_t)k * sk, F * F * 4, bar);
          }
        }
        if (staged) {
          bulk_g2s(smem_u32(st + L.csr), colidx + a_lo, e_bytes, bar);
          bulk_g2s(smem_u32(st + L.csr + kEdgeCap * MT * 4), vals + a_lo, e_bytes, bar);
        }
        if (rp_bytes) bulk_g2s(smem_u32(st + L.csr + kEdgeCap * MT * 8), rowptr + rp_lo, rp_bytes, bar);
      }
      ++t;
      g += cnt;
    }
  }
  // end marker
  const int s = t & 1;
  if (t >= 2) wait(&M->used[s], (uint32_t)(((t >> 1) - 1) & 1));
  if (lane == 0) {
    M->desc[s].cnt = 0;
    M->desc[s].n = 0;
    mbar_arrive(smem_u32(&M->full[s]));
  }
}

// ------------------------------------------------------------------------------------------------
// gather of CPT 16-byte cores: acc[c] = sum_e vals[e] * T[colidx[e] - r0][core q0 + c]  out of a canonical slab
// ------------------------------------------------------------------------------------------------
template <int KC, int CPT, bool STAGED>
__device__ __forceinline__ void gather_cores(float4 (&acc)[CPT], const unsigned char* __restrict__ slab_q, int r0,
                                             const int32_t* __restrict__ ci, const float* __restrict__ cv, int e0,
                                             int e1) {
  float2 a2[2 * CPT];
#pragma unroll
  for (int i = 0; i < 2 * CPT; ++i) a2[i] = make_float2(0.f, 0.f);
  int e = e0;
  for (; e + 1 < e1; e += 2) {   // two edges per trip
    const int cA = (STAGED ? ci[e] : __ldg(ci + e)) - r0, cB = (STAGED ? ci[e + 1] : __ldg(ci + e + 1)) - r0;
    const float wA = STAGED ? cv[e] : __ldg(cv + e), wB = STAGED ? cv[e + 1] : __ldg(cv + e + 1);
    const unsigned char* pA = slab_q + canon_row<KC>(cA);
    const unsigned char* pB = slab_q + canon_row<KC>(cB);
    float4 a[CPT], b[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      a[c] = *reinterpret_cast<const float4*>(pA + c * 128);
      b[c] = *reinterpret_cast<const float4*>(pB + c * 128);
    }
    const float2 wwA = make_float2(wA, wA), wwB = make_float2(wB, wB);
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      a2[2 * c] = __ffma2_rn(wwA, make_float2(a[c].x, a[c].y), a2[2 * c]);
      a2[2 * c + 1] = __ffma2_rn(wwA, make_float2(a[c].z, a[c].w), a2[2 * c + 1]);
      a2[2 * c] = __ffma2_rn(wwB, make_float2(b[c].x, b[c].y), a2[2 * c]);
      a2[2 * c + 1] = __ffma2_rn(wwB, make_float2(b[c].z, b[c].w), a2[2 * c + 1]);
    }
  }
  if (e < e1) {
    const int cA = (STAGED ? ci[e] : __ldg(ci + e)) - r0;
    const float wA = STAGED ? cv[e] : __ldg(cv + e);
    const unsigned char* pA = slab_q + canon_row<KC>(cA);
    const float2 wwA = make_float2(wA, wA);
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const float4 a = *reinterpret_cast<const float4*>(pA + c * 128);
      a2[2 * c] = __ffma2_rn(wwA, make_float2(a.x, a.y), a2[2 * c]);
      a2[2 * c + 1] = __ffma2_rn(wwA, make_float2(a.z, a.w), a2[2 * c + 1]);
    }
  }
#pragma unroll
  for (int c = 0; c < CPT; ++c) acc[c] = make_float4(a2[2 * c].x, a2[2 * c].y, a2[2 * c + 1].x, a2[2 * c + 1].y);
}

// MMA warp: D(m-tile mt)[128, N] (+)= A_hi.B_hi + A_hi.B_lo + A_lo.B_hi over the F/8 k-steps of one order
template <int F, int MT>
__device__ __forceinline__ void issue_order(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                            bool accumulate) {
  constexpr int KC = Geo<F>::KC, N = Geo<F>::N;
  const uint32_t idesc = make_idesc(128, N);
  constexpr uint32_t sbo = KC * 128;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int ks = 0; ks < F / 8; ++ks) {
      const uint32_t ao = (uint32_t)mt * 128 * F * 4 + ks * 256;
      const uint64_t aH = make_desc(a_hi + ao, 128, sbo), aL = make_desc(a_lo + ao, 128, sbo);
      const uint64_t bH = make_desc(b_hi + ks * 256, 128, sbo), bL = make_desc(b_lo + ks * 256, 128, sbo);
      const uint32_t d = tmem_d + (uint32_t)mt * N;
      mma_ss(d, aH, bH, idesc, (accumulate || ks > 0) ? 1u : 0u);
      mma_ss(d, aH, bL, idesc, 1u);
      mma_ss(d, aL, bH, idesc, 1u);
    }
  }
}

__device__ __forceinline__ int local_graph(const TileDesc& d, int grow) {
  int j = 0;
#pragma unroll
  for (int i = 1; i < 9; ++i) j += (i < d.cnt && grow >= d.gp[i]) ? 1 : 0;
  return j;
}

__device__ __forceinline__ float4 ld4s(const unsigned char* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4s(unsigned char* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

template <int CPT>
__device__ __forceinline__ void tmem_ld_cores(uint32_t taddr, float4 (&v)[CPT]) {
  static_assert(CPT == 1 || CPT == 2 || CPT == 4, "cores per thread");
  if constexpr (CPT == 4) {
    float f[16];
    tmem_ld16(taddr, f);
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
  } else if constexpr (CPT == 1) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
    wait_ld();
    v[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
  } else {
    float f[8];
    tmem_ld8(taddr, f);
    v[0] = make_float4(f[0], f[1], f[2], f[3]);
    v[1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// thread geometry: a compute warp owns 32 rows (rg) x CPT consecutive 16-byte cores (first core q0)
template <int F, int MT>
struct Map {
  static constexpr int KC = F / 4;
#ifndef FETA_TILE_CPT16
#define FETA_TILE_CPT16 4
#endif
  // cores per thread: F = 16 -> FETA_TILE_CPT16 (4: lane owns the whole row), F = 8 -> 2 (whole row)
  static constexpr int CPT = F == 16 ? FETA_TILE_CPT16 : 2;
  static constexpr int NCW = 4 * MT * KC / CPT;                // compute warps
  static constexpr int NT = 32 * NCW;                          // compute threads
  static constexpr int THREADS = NT + 64;                      // + MMA warp + loader warp
  static constexpr int MINB = MT == 2 ? 1 : (F == 16 ? 2 : 3);  // CTAs per SM the register budget must allow
  static_assert(NT <= 960, "block size");
};

// =====================================================================================================
// forward
// =====================================================================================================
template <int F, int MT>
__global__ void __launch_bounds__(Map<F, MT>::THREADS, Map<F, MT>::MINB) cheb_fwd_tile_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
    const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr, const int32_t* __restrict__ row_graph,
    const float* __restrict__ theta, int64_t sk, int64_t sg, const float* __restrict__ bias, float* __restrict__ out,
    int64_t R, int64_t G, int K, int window, int32_t* meta, int max_nodes) {
  using MP = Map<F, MT>;
  constexpr int ROWS = 128 * MT, KC = Geo<F>::KC, N = Geo<F>::N, CPT = MP::CPT, NT = MP::NT;
  constexpr uint32_t SLAB = ROWS * F * 4;
  extern __shared__ __align__(1024) unsigned char smem[];
  if (!plan_guard_ok(meta, G, max_nodes)) { nan_fill(out, R * F); return; }
  const Smem L = carve<F, MT>(K, 2, 1, true, false, true, 0);
  Misc* M = reinterpret_cast<Misc*>(smem + L.misc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int W_MMA = MP::NCW, W_LOAD = MP::NCW + 1;
  constexpr uint32_t TCOLS = 2 * MT * N;                // two accumulator sets: the epilogue of tile t is deferred

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&M->full[i]), 1);
      mbar_init(smem_u32(&M->used[i]), MP::NCW);          // one arrival per compute WARP: an mbarrier arrive
      mbar_init(smem_u32(&M->a_ready[i]), MP::NCW);       // is a serialised shared-memory atomic

    }
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&M->mma[i]), 1);
    mbar_fence_init();
  }
  if (tid < F) M->bias[tid] = bias ? __ldg(bias + tid) : 0.0f;
  if (warp == W_MMA) tmem_alloc(smem_u32(&M->tmem), TCOLS);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = M->tmem;

  if (warp == W_LOAD) {
    loader_warp<F, MT>(smem, L, M, x, nullptr, theta, sk, sg, rowptr, colidx, vals, graph_ptr, row_graph, R, G, K,
                       window);
  } else if (warp == W_MMA) {
    // ---------------- MMA issuer
    for (int t = 0;; ++t) {
      wait(&M->full[t & 1], (uint32_t)((t >> 1) & 1));
      if (M->desc[t & 1].cnt == 0) break;
      for (int k = 0; k < K; ++k) {
        const int o = t * K + k;
        // T_k (hi, lo) and -- k = 0 -- the split Theta are in place.  Two alternating mbarriers: the compute
        // threads can never be two orders ahead of this warp (they wait on the commits below)
        wait_short(&M->a_ready[o & 1], (uint32_t)((o >> 1) & 1));
        fence_after();
        if (lane == 0) {
          issue_order<F, MT>(tmem + (uint32_t)(t & 1) * MT * N, smem_u32(smem + L.ahi + (uint32_t)(k & 1) * SLAB),
                             smem_u32(smem + L.alo), smem_u32(smem + L.bhi + (uint32_t)k * N * F * 4),
                             smem_u32(smem + L.blo + (uint32_t)k * N * F * 4), k > 0);
          mma_commit(smem_u32(&M->mma[o & 3]));
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- compute threads: warp = (32-row group rg, cores q0 .. q0+CPT-1), lane = row
    const int rg = warp % (4 * MT), q0 = (warp / (4 * MT)) * CPT;
    const int row = rg * 32 + lane;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32)) << 16;
    const uint32_t my = canon_row<KC>(row) + q0 * 128;
    // Theta re-lay: per-thread constants (one division per kernel, none per tile)
    constexpr int TH_PER = F * KC, TH_STEP = NT / TH_PER;
    const int th_jk0 = tid / TH_PER, th_j0 = th_jk0 / K, th_k0 = th_jk0 - th_j0 * K;
    const int th_sdiv = TH_STEP / K, th_smod = TH_STEP - th_sdiv * K;
    const int th_o = tid % F, th_q = (tid / F) % KC;
    const int th_src0 = (4 * th_q) * F + th_o;
    const uint32_t th_dst0 = canon_row<KC>(th_o) + th_q * 128;
    // deferred epilogue state of the previous tile
    bool ep_pending = false, ep_active = false;
    int ep_row = 0, ep_j = 0, ep_jlo = 0, ep_jhi = -1, ep_o = 0;
    uint32_t ep_d = 0;

    auto epilogue = [&]() {
      wait_short(&M->mma[ep_o & 3], (uint32_t)((ep_o >> 2) & 1));
      fence_after();
      float4 acc[CPT];
#pragma unroll
      for (int c = 0; c < CPT; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int jj = ep_jlo; jj <= ep_jhi; ++jj) {        // warp-uniform bounds: the graphs present in this warp
        float4 v[CPT];
        tmem_ld_cores<CPT>(ep_d + (uint32_t)jj * F, v);
        if (jj == ep_j) {
#pragma unroll
          for (int c = 0; c < CPT; ++c) acc[c] = v[c];
        }
      }
      if (ep_active) {
        float* orow = out + (size_t)ep_row * F + 4 * q0;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          const float4 b = *reinterpret_cast<const float4*>(&M->bias[4 * (q0 + c)]);
          *reinterpret_cast<float4*>(orow + 4 * c) =
              make_float4(acc[c].x + b.x, acc[c].y + b.y, acc[c].z + b.z, acc[c].w + b.w);
        }
      }
      fence_before();          // these tcgen05.ld are ordered before the MMAs that reuse the accumulator set
      ep_pending = false;
    };

    for (int t = 0;; ++t) {
      const int s = t & 1;
      wait(&M->full[s], (uint32_t)((t >> 1) & 1));
      const TileDesc d = M->desc[s];
      if (d.cnt == 0) break;
      const unsigned char* st = smem + L.stage + (size_t)s * L.stage_bytes;
      const bool active = row < d.n;
      const int32_t* rp = reinterpret_cast<const int32_t*>(st + L.csr + kEdgeCap * MT * 8) + (d.r0 - d.rp_base);
      const int e0 = active ? rp[row] : 0, e1 = active ? rp[row + 1] : 0;
      float4 p1[CPT], p2[CPT];                 // own cores of T_{k-1}, T_{k-2}
      // ---- phase 0: x -> T_0 (hi = raw, lo), Theta -> transposed K-major hi / lo blocks
      if (t > 0) {   // every MMA of tile t-1 has completed (lo slab, B blocks and hi slabs are free again)
        const int o = t * K - 1;
        wait_short(&M->mma[o & 3], (uint32_t)((o >> 2) & 1));
      }
      if (active) {
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
          p1[c] = ld4s(st + L.xd + (uint32_t)row * F * 4 + (q0 + c) * 16);
          st4s(smem + L.ahi + my + c * 128, p1[c]);
          st4s(smem + L.alo + my + c * 128, tf32_lo4(p1[c]));
        }
      }
      {
        // B_k[n = j*F + o][kk = i] = Theta_k(g_j)[i][o]: a thread owns one (o, q = i/4) and walks the (graph, order)
        // blocks; canon_row(j*F + o) = j*F*F*4 + canon_row(o) for F in {8, 16}
        const int njk = d.cnt * K;
        int jk = th_jk0, j = th_j0, k = th_k0;
        const float* src = reinterpret_cast<const float*>(st + L.th) + (size_t)jk * F * F + th_src0;
        for (; jk < njk; jk += TH_STEP) {
          const float4 v = make_float4(src[0], src[F], src[2 * F], src[3 * F]);
          const uint32_t off = (uint32_t)k * (N * F * 4) + (uint32_t)j * (F * F * 4) + th_dst0;
          st4s(smem + L.bhi + off, v);
          st4s(smem + L.blo + off, tf32_lo4(v));
          src += TH_STEP * F * F;
          k += th_smod;
          j += th_sdiv;
          if (k >= K) { k -= K; ++j; }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&M->a_ready[(t * K) & 1]));
      // the previous tile's epilogue runs here: its last MMAs had the whole phase 0 to finish
      if (ep_pending) epilogue();
      // epilogue bookkeeping of THIS tile (the descriptor slot is recycled before the epilogue runs)
      {
        const int j = local_graph(d, d.r0 + row);
        int jlo = active ? j : 8, jhi = active ? j : -1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          jlo = min(jlo, __shfl_xor_sync(0xffffffffu, jlo, off));
          jhi = max(jhi, __shfl_xor_sync(0xffffffffu, jhi, off));
        }
        ep_pending = true; ep_active = active; ep_row = d.r0 + row; ep_j = j; ep_jlo = jlo; ep_jhi = jhi;
        ep_o = t * K + K - 1;
        ep_d = tmem + lane_base + (uint32_t)(t & 1) * MT * N + (uint32_t)(row >> 7) * N + 4 * q0;
      }
      bar_sync(1, NT);
      // ---- orders 1 .. K-1:  T_k = c_k L T_{k-1} - T_{k-2}  (own cores of T_{k-1}, T_{k-2} stay in registers)
      for (int k = 1; k < K; ++k) {
        const unsigned char* prev = smem + L.ahi + (uint32_t)((k - 1) & 1) * SLAB + q0 * 128;
        unsigned char* dst = smem + L.ahi + (uint32_t)(k & 1) * SLAB + my;
        float4 cur[CPT];
        if (active) {
          // the staged slice stays typed as a shared-memory pointer (a select against the global arrays would
          // turn every access into a generic load)
          if (d.e_base >= 0)
            gather_cores<KC, CPT, true>(cur, prev, d.r0, reinterpret_cast<const int32_t*>(st + L.csr),
                                        reinterpret_cast<const float*>(st + L.csr + kEdgeCap * MT * 4),
                                        e0 - d.e_base, e1 - d.e_base);
          else gather_cores<KC, CPT, false>(cur, prev, d.r0, colidx, vals, e0, e1);
          if (k >= 2) {
#pragma unroll
            for (int c = 0; c < CPT; ++c)
              cur[c] = make_float4(fmaf(2.0f, cur[c].x, -p2[c].x), fmaf(2.0f, cur[c].y, -p2[c].y),
                                   fmaf(2.0f, cur[c].z, -p2[c].z), fmaf(2.0f, cur[c].w, -p2[c].w));
          }
#pragma unroll
          for (int c = 0; c < CPT; ++c) st4s(dst + c * 128, cur[c]);   // hi slab k&1: MMA(k-2) finished (below)
        }
        {   // the single lo slab was read by MMA(k-1)
          const int o = t * K + k - 1;
          wait_short(&M->mma[o & 3], (uint32_t)((o >> 2) & 1));
        }
        if (active) {
#pragma unroll
          for (int c = 0; c < CPT; ++c) {
            st4s(smem + L.alo + my + c * 128, tf32_lo4(cur[c]));
            p2[c] = p1[c];
            p1[c] = cur[c];
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&M->a_ready[(t * K + k) & 1]));
        bar_sync(1, NT);
      }
      if (lane == 0) mbar_arrive(smem_u32(&M->used[s]));     // landing stage consumed (all lanes passed bar_sync)
    }
    if (ep_pending) epilogue();
  }
  fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, TCOLS);
}

// ------------------------------------------------------------------------------------------------
static inline int pick_window(int64_t R, int rows, int ctas) {
  // rows per persistent work item: large enough that the tail tile of a window is rare, small enough that
  // every CTA gets work on small batches
  int64_t w = R / ((int64_t)ctas * 2);
  if (w > 2048) w = 2048;
  if (w < rows) w = rows;
  return (int)w;
}

template <int F, int MT>
static int launch_fwd_tile(const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                           const int32_t* graph_ptr, const int32_t* row_graph, const float* theta, int64_t sk, int64_t sg,
                           const float* bias, float* out, int64_t R, int64_t G, int K, int32_t* meta, int max_nodes,
                           cudaStream_t st) {
  const Smem L = carve<F, MT>(K, 2, 1, true, false, true, 0);
  if (L.total > 227 * 1024) return 1;
  auto kern = cheb_fwd_tile_kernel<F, MT>;
  FETA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  int per_sm = (int)((227 * 1024) / (L.total + 1024));
  const int by_threads = 2048 / Map<F, MT>::THREADS, by_tmem = 512 / (2 * MT * Geo<F>::N);
  if (per_sm > by_threads) per_sm = by_threads;
  if (per_sm > by_tmem) per_sm = by_tmem;
  if (per_sm < 1) per_sm = 1;
  const int ctas = kNumSMs * per_sm;
  const int window = pick_window(R, 128 * MT, ctas);
  int64_t grid = ceil_div(R, window);
  if (grid > ctas) grid = ctas;
  kern<<<(unsigned)grid, Map<F, MT>::THREADS, L.total, st>>>(x, rowptr, colidx, vals, graph_ptr, row_graph, theta, sk,
                                                           sg, bias, out, R, G, K, window, meta, max_nodes);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

}  // namespace tile

// returns FETA_OK if launched, 1 if the shape is not eligible (the caller falls back), < 0 on error
int cheb_fwd_tile_try(const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                      const int32_t* graph_ptr, const int32_t* row_graph, const float* theta, int64_t sk, int64_t sg,
                      const float* bias, float* out, int64_t R, int64_t G, int K, int F, int max_nodes, int32_t* meta,
                      cudaStream_t st) {
  // Measured (profiles/r2_cheb_tile.md): correct everywhere, but on <= 64-row graphs the warp-per-graph kernel is
  // faster (preparing hi/lo operands for the 3xTF32 tcgen05 products costs as many issue slots as the mma.sync
  // fragments it replaces, and ~110 KB of shared memory per tile caps residency at 8 compute warps per SM).
  // Default: only where no better kernel exists (graphs of 65..128 rows in molecule-size batches);
  // FETA_CHEB_TILE=1 forces it for every eligible shape, FETA_CHEB_NO_TILE_KERNEL=1 disables it.
  if (getenv("FETA_CHEB_NO_TILE_KERNEL") != nullptr) return 1;
  if (getenv("FETA_CHEB_TILE") == nullptr && !(max_nodes > 64 && max_nodes <= 128)) return 1;
  if (!(F == 8 || F == 16) || K < 1 || max_nodes < 1 || max_nodes > 256) return 1;
  if ((size_t)K * F * F * 4 * (64 / F) > 16 * 1024) return 1;
  if (((uintptr_t)colidx % 16) || ((uintptr_t)vals % 16) || ((uintptr_t)rowptr % 16)) return 1;
#define FETA_TILE_CASE(F_, MT_)                                                                                     \
  if (F == F_ && ((max_nodes <= 128) ? 1 : 2) == MT_)                                                                \
    return tile::launch_fwd_tile<F_, MT_>(x, rowptr, colidx, vals, graph_ptr, row_graph, theta, sk, sg, bias, out, R, \
                                          G, K, meta, max_nodes, st);
  FETA_TILE_CASE(8, 1) FETA_TILE_CASE(8, 2) FETA_TILE_CASE(16, 1) FETA_TILE_CASE(16, 2)
#undef FETA_TILE_CASE
  return 1;
}

}  // namespace feta
