// A1/A3: ChebConvDynamic forward for batches of small graphs (every graph <= 64 rows) -- second generation of the
// warp-per-graph kernel (csrc/cheb_warp.cu), written against its ncu source page (profiles/r2_cheb_lane.md):
// cheb_warp.cu executed 1686 warp instructions per 23-row graph of which only 98 were MMAs and 190 shared-memory
// loads of payload -- 19 % IMAD and 12 % LOP3 were address arithmetic on run-time buffer geometry, the gather loop
// re-read and re-decoded the CSR entries of a row in every Chebyshev order, and nothing was unrolled across orders.
// Here
//   * K, F and the rows-per-lane count are template parameters: the whole recursion of a graph is straight-line
//     code, every shared-memory access is [per-lane register + immediate];
//   * the first four CSR entries of a row are decoded ONCE per graph into (shared-memory address of the neighbour
//     row, weight) register pairs -- a molecule row has <= 4 neighbours -- and every order replays them as
//     4 x (LDS.128 x F/4, FFMA2 x F/2); longer rows finish in a generic loop;
//   * T_k . Theta_k runs on the tensor cores for F = 8 as well as F = 16 (m16n8k8 TF32, both operands split
//     hi + lo, three MMAs per product: fp32-grade), MMAs of order k and the gather of order k+1 are adjacent in
//     program order with no barrier between them (both only read T_k), so the scheduler overlaps them;
//   * the x slab lands by cp.async.bulk and is re-laid into the XOR-swizzled layout by its own row's lane, in
//     place, with no extra synchronisation (the permutation stays inside a 16-float row).
// The persistent structure is cheb_warp.cu's: one warp = one graph at a time (lane = row), TMA copies issued one
// graph ahead on mbarriers, the scalars that size them two and three graphs ahead, one Theta buffer refilled
// behind the epilogue, the output slab returned by one bulk store.
#include <stdlib.h>

#include "common.cuh"
#include "graph_tile.cuh"
#include "umma.cuh"

namespace feta {
namespace lane {

using namespace tc;

__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Shared memory is addressed through 32-bit shared-window addresses computed by hand (ld/st.shared): the XOR
// swizzle defeats the compiler's [base + immediate] folding, and generic pointers cost an IADD3 per access.
// The "memory" clobber orders them against each other and against __syncwarp at the compiler level; ptxas still
// schedules the loads freely.
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float2 lds64(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, float2 v) {
  asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(a), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ float lds32(uint32_t a) {     // constant offsets fold into the addressing mode in ptxas
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ int32_t lds32i(uint32_t a) {
  int32_t v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}

// swizzle of a [rows, F] fp32 slab: 16-byte chunk q of row r is stored at chunk q ^ swz<F>(r).  Lane-per-row
// LDS.128 / STS.128 and the m16n8k8 A-fragment loads (8 rows x 4 consecutive words) are both conflict-free.
template <int F>
__device__ __forceinline__ uint32_t swz(uint32_t r) {
  return F == 16 ? ((r >> 1) & 3u) : ((r >> 2) & 1u);
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  // the tensor core reads the top 19 bits of a tf32 operand: "hi" is x itself, the remainder is exact in fp32
  hi = __float_as_uint(x);
  lo = __float_as_uint(x - __uint_as_float(hi & 0xFFFFE000u));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

struct Cfg {
  int rpl, warps, nnz_cap, rows_cap;
  uint32_t per_warp;
  size_t smem;
  bool ok;
};

constexpr uint32_t kBarBytes = 64;

static inline uint32_t csr_bytes_of(int nnz_cap) { return (uint32_t)align_up((size_t)(nnz_cap + 8) * 4, 64); }

static Cfg config(int F, int K, int max_nodes) {
  Cfg c{0, 0, 0, 0, 0, 0, false};
  if (!(F == 8 || F == 16) || K < 1 || K > 4 || max_nodes < 1 || max_nodes > 64) return c;
  c.rpl = max_nodes <= 32 ? 1 : 2;
  c.rows_cap = (max_nodes + 15) / 16 * 16;       // whole m16 tiles
  c.nnz_cap = c.rows_cap * 4;
  const size_t slab = (size_t)c.rows_cap * F * 4;
  const size_t stage = slab + 2 * csr_bytes_of(c.nnz_cap);
  c.per_warp = (uint32_t)align_up(kBarBytes + slab + (size_t)K * F * F * 4 + 2 * stage, 128);
  int w = (int)((227 * 1024) / c.per_warp);
  if (w > (c.rpl == 1 ? 16 : 12)) w = c.rpl == 1 ? 16 : 12;
  if (w < 4) return c;
  c.warps = w;
  c.smem = (size_t)c.per_warp * w;
  c.ok = true;
  return c;
}

template <int RPL>
struct GraphDesc {  // scalars of one graph, fetched ahead of use
  int r0, r1, e_lo, e_hi;
  int e0[RPL], e1[RPL];
};

// acc[mt][nt] += T[16mt .. 16mt+15, :] . Theta_k  (3xTF32).  `ta[j]` = shared-memory byte address of this lane's
// A-fragment word in 16-byte chunk j of row g (rows 16mt + g, + 8 are immediates); `tb` = address of
// Theta_k[tq][g].
template <int F, int MTMAX>
__device__ __forceinline__ void mma_order(float (&acc)[MTMAX][F / 8][4], const uint32_t (&ta)[F / 4], uint32_t tb,
                                          int MT) {
  constexpr int NT = F / 8, KS = F / 8;
  uint32_t bh[NT][KS][2], bl[NT][KS][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      split_tf32(lds32(tb + ((8 * ks) * F + 8 * nt) * 4), bh[nt][ks][0], bl[nt][ks][0]);
      split_tf32(lds32(tb + ((8 * ks + 4) * F + 8 * nt) * 4), bh[nt][ks][1], bl[nt][ks][1]);
    }
#pragma unroll
  for (int mt = 0; mt < MTMAX; ++mt) {
    if (mt < MT) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ah[4], al[4];
        split_tf32(lds32(ta[2 * ks] + (16 * mt) * F * 4), ah[0], al[0]);
        split_tf32(lds32(ta[2 * ks] + (16 * mt + 8) * F * 4), ah[1], al[1]);
        split_tf32(lds32(ta[2 * ks + 1] + (16 * mt) * F * 4), ah[2], al[2]);
        split_tf32(lds32(ta[2 * ks + 1] + (16 * mt + 8) * F * 4), ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[mt][nt], ah, bh[nt][ks]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[mt][nt], ah, bl[nt][ks]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[mt][nt], al, bh[nt][ks]);
      }
    }
  }
}

// Same product against Theta_k^T (the dx kernel: dx = sum_k T_k(L^T) dOut . Theta_k^T).  The contraction index of an
// MMA may be permuted freely as long as A and B agree; here lane (g, tq) owns the CONTIGUOUS indices
// F/4 * tq .. F/4 * tq + F/4 - 1 (k-step ks, half h -> F/4 * tq + 2 ks + h), so its A values of a row are one
// vector load of that row's slab and its B values one vector load of row g of Theta_k -- both conflict-free
// (a scalar B-fragment load of Theta^T would be 8-way bank conflicted).  `ta` = address of (row g, first own
// channel) in the slab, `tb` = address of Theta_k[g][first own channel].
template <int F, int MTMAX>
__device__ __forceinline__ void mma_order_t(float (&acc)[MTMAX][F / 8][4], uint32_t ta, uint32_t tb, int MT) {
  constexpr int NT = F / 8, KS = F / 8, V = F / 4;   // V own contraction indices per lane
  uint32_t bh[NT][KS][2], bl[NT][KS][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    float v[V];
    if constexpr (F == 16) {
      const float4 t = lds128(tb + (8 * nt) * F * 4);
      v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    } else {
      const float2 t = lds64(tb + (8 * nt) * F * 4);
      v[0] = t.x, v[1] = t.y;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      split_tf32(v[2 * ks], bh[nt][ks][0], bl[nt][ks][0]);
      split_tf32(v[2 * ks + 1], bh[nt][ks][1], bl[nt][ks][1]);
    }
  }
#pragma unroll
  for (int mt = 0; mt < MTMAX; ++mt) {
    if (mt < MT) {
      float r0[V], r1[V];
      if constexpr (F == 16) {
        const float4 t0 = lds128(ta + (16 * mt) * F * 4), t1 = lds128(ta + (16 * mt + 8) * F * 4);
        r0[0] = t0.x, r0[1] = t0.y, r0[2] = t0.z, r0[3] = t0.w;
        r1[0] = t1.x, r1[1] = t1.y, r1[2] = t1.z, r1[3] = t1.w;
      } else {
        const float2 t0 = lds64(ta + (16 * mt) * F * 4), t1 = lds64(ta + (16 * mt + 8) * F * 4);
        r0[0] = t0.x, r0[1] = t0.y;
        r1[0] = t1.x, r1[1] = t1.y;
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ah[4], al[4];
        split_tf32(r0[2 * ks], ah[0], al[0]);
        split_tf32(r1[2 * ks], ah[1], al[1]);
        split_tf32(r0[2 * ks + 1], ah[2], al[2]);
        split_tf32(r1[2 * ks + 1], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[mt][nt], ah, bh[nt][ks]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[mt][nt], ah, bl[nt][ks]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[mt][nt], al, bh[nt][ks]);
      }
    }
  }
}

constexpr int kSlots = 4;   // CSR entries of a row decoded into registers once per graph

// TRANS = false: out = sum_k T_k(L) x . Theta_k + bias (forward).  TRANS = true: the same recursion against
// Theta_k^T -- called with (dOut, the source-grouped CSR = L^T, no bias) it is the input gradient, because the
// operator acts on rows and the filters on channels: dx = sum_k T_k(L^T) dOut . Theta_k^T.
template <int F, int RPL, int K, bool TRANS>
__global__ void __launch_bounds__(RPL == 1 ? 512 : 384, 1) cheb_fwd_lane_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
    const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr, const float* __restrict__ theta,
    int64_t sk, int64_t sg, const float* __restrict__ bias, float* __restrict__ out, int64_t R, int64_t G,
    int nnz_cap, int rows_cap, uint32_t per_warp_bytes, int32_t* meta, int max_nodes) {
  constexpr int Q = F / 4, NT = F / 8, MTMAX = 2 * RPL;
  constexpr uint32_t ROWB = F * 4;
  extern __shared__ __align__(128) unsigned char sm[];
  if (!plan_guard_ok(meta, G, max_nodes)) { nan_fill(out, R * F); return; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * warps + warp, stride = (int64_t)gridDim.x * warps;
  if (gw >= G) return;
  const int n_it = (int)((G - gw + stride - 1) / stride);

  // ---- per-warp shared-memory carve: 32-bit shared-window addresses, every buffer 64-byte aligned
  const uint32_t slab = (uint32_t)rows_cap * ROWB;
  const uint32_t csr_bytes = (uint32_t)(((nnz_cap + 8) * 4 + 63) / 64 * 64);
  constexpr uint32_t th_bytes = (uint32_t)K * F * F * 4;
  const uint32_t base = smem_u32(sm) + (uint32_t)warp * per_warp_bytes;
  const uint32_t a_bar = base, a_odd = base + kBarBytes, a_th = a_odd + slab, a_st0 = a_th + th_bytes;
  const uint32_t stage_bytes = slab + 2 * csr_bytes;
  const bool theta_contig = (sk == (int64_t)F * F);

  if (lane == 0) {
    mbar_init(a_bar, 1);
    mbar_init(a_bar + 8, 1);
    mbar_init(a_bar + 16, 1);
    mbar_fence_init();
  }
  __syncwarp();
  const int nnz_total = __ldg(rowptr + R);

  // ---- lane constants
  const uint32_t g = (uint32_t)lane >> 2, tq = (uint32_t)lane & 3u;
  const uint32_t sw_row = swz<F>((uint32_t)lane);               // rows lane and lane + 32 share it
  uint32_t own[Q];                                               // own row, chunk q: byte offset inside a slab
#pragma unroll
  for (int q = 0; q < Q; ++q) own[q] = (uint32_t)lane * ROWB + (((uint32_t)q ^ sw_row) << 4);
  uint32_t afr[Q];                                               // A fragment: row g, chunk j, word tq
#pragma unroll
  for (int j = 0; j < Q; ++j) afr[j] = g * ROWB + (((uint32_t)j ^ swz<F>(g)) << 4) + tq * 4;
  // TRANS: one vector of own contraction indices (see mma_order_t)
  const uint32_t afr_t = F == 16 ? g * ROWB + ((tq ^ swz<F>(g)) << 4) : g * ROWB + (((tq >> 1) ^ swz<F>(g)) << 4) + (tq & 1u) * 8;
  const uint32_t bfr = TRANS ? a_th + (g * F + (F / 4) * tq) * 4   // Theta_k[g][own channels]
                             : a_th + (tq * F + g) * 4;            // Theta_k[tq][g] (+ k F F 4 + immediates)
  const uint32_t ofr = a_odd + (g * F + 2 * tq) * 4;             // output staging (dense) in the odd buffer
  uint32_t own_o[Q], ta_o[Q];                                    // the odd buffer never moves
#pragma unroll
  for (int q = 0; q < Q; ++q) own_o[q] = a_odd + own[q], ta_o[q] = a_odd + afr[q];
  float bias_m[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    bias_m[nt][0] = bias ? __ldg(bias + 8 * nt + 2 * tq) : 0.0f;
    bias_m[nt][1] = bias ? __ldg(bias + 8 * nt + 2 * tq + 1) : 0.0f;
  }

  auto fetch_a = [&](int it, GraphDesc<RPL>& d) {
    d.r0 = d.r1 = 0;
    if (it < n_it) {
      const int64_t gi = gw + (int64_t)it * stride;
      d.r0 = __ldg(graph_ptr + gi);
      d.r1 = __ldg(graph_ptr + gi + 1);
    }
  };
  auto fetch_b = [&](int it, GraphDesc<RPL>& d) {
    d.e_lo = d.e_hi = 0;
#pragma unroll
    for (int m = 0; m < RPL; ++m) d.e0[m] = d.e1[m] = 0;
    if (it < n_it) {
      d.e_lo = __ldg(rowptr + d.r0);
      d.e_hi = __ldg(rowptr + d.r1);
#pragma unroll
      for (int m = 0; m < RPL; ++m) {
        const int row = lane + 32 * m;
        if (row < d.r1 - d.r0) {
          d.e0[m] = __ldg(rowptr + d.r0 + row);
          d.e1[m] = __ldg(rowptr + d.r0 + row + 1);
        }
      }
    }
  };
  auto staged_csr = [&](const GraphDesc<RPL>& d, int& a_lo, int& a_hi) {
    a_lo = d.e_lo & ~3;
    a_hi = (d.e_hi + 3) & ~3;
    return (a_hi - a_lo <= nnz_cap) && (a_hi <= nnz_total);
  };
  auto issue = [&](int it, const GraphDesc<RPL>& d) {   // TMA copies of graph `it` into stage it & 1
    if (it >= n_it) return;
    const uint32_t st = a_st0 + (uint32_t)(it & 1) * stage_bytes;
    const uint32_t bar = a_bar + 8u * (uint32_t)(it & 1);
    const int n = d.r1 - d.r0;
    int a_lo, a_hi;
    const bool copy_csr = staged_csr(d, a_lo, a_hi) && (d.e_hi > d.e_lo);
    const uint32_t bytes = (uint32_t)n * ROWB + (copy_csr ? 2u * (uint32_t)(a_hi - a_lo) * 4u : 0u);
    fence_proxy_async();   // generic-proxy writes into this stage (even orders) before the TMA refills it
    __syncwarp();
    if (lane == 0) {
      mbar_arrive_expect_tx(bar, bytes);
      if (n > 0) bulk_g2s(st, x + (size_t)d.r0 * F, (uint32_t)n * ROWB, bar);
      if (copy_csr) {
        bulk_g2s(st + slab, colidx + a_lo, (uint32_t)(a_hi - a_lo) * 4, bar);
        bulk_g2s(st + slab + csr_bytes, vals + a_lo, (uint32_t)(a_hi - a_lo) * 4, bar);
      }
    }
  };
  auto issue_theta = [&](int it) {
    if (it >= n_it) return;
    if (lane == 0) {
      const int64_t gi = gw + (int64_t)it * stride;
      const uint32_t bar = a_bar + 16;
      mbar_arrive_expect_tx(bar, th_bytes);
      if (theta_contig) {
        bulk_g2s(a_th, theta + gi * sg, th_bytes, bar);
      } else {
        for (int k = 0; k < K; ++k)
          bulk_g2s(a_th + (uint32_t)k * F * F * 4, theta + gi * sg + (int64_t)k * sk, F * F * 4, bar);
      }
    }
  };

  GraphDesc<RPL> d0, d1, d2, d3;
  fetch_a(0, d0);
  fetch_a(1, d1);
  fetch_a(2, d2);
  fetch_b(0, d0);
  fetch_b(1, d1);
  issue(0, d0);
  issue_theta(0);

  for (int it = 0; it < n_it; ++it) {
    fetch_a(it + 3, d3);
    fetch_b(it + 2, d2);
    issue(it + 1, d1);

    const uint32_t a_even = a_st0 + (uint32_t)(it & 1) * stage_bytes;   // x slab = T_0, later T_2
    const uint32_t a_ci = a_even + slab, a_cv = a_ci + csr_bytes;
    const int n = d0.r1 - d0.r0;
    const int MT = (n + 15) >> 4;
    int a_lo, a_hi;
    const bool staged = staged_csr(d0, a_lo, a_hi);
    uint32_t own_e[Q], ta_e[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) own_e[q] = a_even + own[q], ta_e[q] = a_even + afr[q];
    mbar_wait_parity(a_bar + 8u * (uint32_t)(it & 1), (uint32_t)((it >> 1) & 1));
    if (lane == 0) bulk_wait_read0();   // the previous graph's output store has drained the odd buffer
    __syncwarp();

    // ---- decode the first kSlots CSR entries of every own row: neighbour-row offset (chunk 0, swizzled) + weight
    uint32_t nb[RPL][kSlots];
    float wt[RPL][kSlots];
    int deg[RPL], nslot[RPL];
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
      deg[m] = d0.e1[m] - d0.e0[m];                  // 0 for lanes past the graph
      nslot[m] = staged ? min(deg[m], kSlots) : 0;
      const uint32_t el = (uint32_t)(d0.e0[m] - a_lo) * 4;
#pragma unroll
      for (int j = 0; j < kSlots; ++j) {
        nb[m][j] = 0;
        wt[m][j] = 0.0f;
      }
      if (nslot[m] > 0) {   // up to 3 entries past the row are read (inside the staged slice + slack), never used
        const uint32_t c0 = (uint32_t)(lds32i(a_ci + el + 0) - d0.r0), c1 = (uint32_t)(lds32i(a_ci + el + 4) - d0.r0);
        const uint32_t c2 = (uint32_t)(lds32i(a_ci + el + 8) - d0.r0), c3 = (uint32_t)(lds32i(a_ci + el + 12) - d0.r0);
        wt[m][0] = lds32(a_cv + el + 0), wt[m][1] = lds32(a_cv + el + 4);
        wt[m][2] = lds32(a_cv + el + 8), wt[m][3] = lds32(a_cv + el + 12);
        nb[m][0] = c0 * ROWB + (swz<F>(c0) << 4), nb[m][1] = c1 * ROWB + (swz<F>(c1) << 4);
        nb[m][2] = c2 * ROWB + (swz<F>(c2) << 4), nb[m][3] = c3 * ROWB + (swz<F>(c3) << 4);
      }
    }

    // ---- T_0: re-lay the own row(s) of the dense x slab into the swizzled layout, in place
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
      if (lane + 32 * m < n) {
        float4 v[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) v[q] = lds128(a_even + (uint32_t)(lane + 32 * m) * ROWB + 16 * q);
#pragma unroll
        for (int q = 0; q < Q; ++q) sts128(own_e[q] + 32 * m * ROWB, v[q]);
      }
    }
    __syncwarp();

    float acc[MTMAX][NT][4];
#pragma unroll
    for (int mt = 0; mt < MTMAX; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.0f;

#pragma unroll
    for (int k = 1; k < K; ++k) {
      const uint32_t a_src = (k & 1) ? a_even : a_odd;   // T_{k-1}
#pragma unroll
      for (int m = 0; m < RPL; ++m) {
        if (lane + 32 * m < n) {
          float2 a2[F / 2];
#pragma unroll
          for (int i = 0; i < F / 2; ++i) a2[i] = make_float2(0.f, 0.f);
#pragma unroll
          for (int j = 0; j < kSlots; ++j) {
            if (j < nslot[m]) {
              const uint32_t p = a_src + nb[m][j];    // buffers are 64-byte aligned: the XOR stays inside the row
              const float2 ww = make_float2(wt[m][j], wt[m][j]);
              float4 a[Q];
#pragma unroll
              for (int q = 0; q < Q; ++q) a[q] = lds128(p ^ ((uint32_t)q << 4));
#pragma unroll
              for (int q = 0; q < Q; ++q) {
                a2[2 * q] = __ffma2_rn(ww, make_float2(a[q].x, a[q].y), a2[2 * q]);
                a2[2 * q + 1] = __ffma2_rn(ww, make_float2(a[q].z, a[q].w), a2[2 * q + 1]);
              }
            }
          }
          if (deg[m] > nslot[m]) {   // rows with more than kSlots neighbours, or a CSR slice that was not staged
            for (int e = d0.e0[m] + nslot[m]; e < d0.e1[m]; ++e) {
              uint32_t c;
              float w;
              if (staged) {
                c = (uint32_t)(lds32i(a_ci + (uint32_t)(e - a_lo) * 4) - d0.r0);
                w = lds32(a_cv + (uint32_t)(e - a_lo) * 4);
              } else {
                c = (uint32_t)(__ldg(colidx + e) - d0.r0);
                w = __ldg(vals + e);
              }
              const uint32_t p = a_src + c * ROWB + (swz<F>(c) << 4);
              const float2 ww = make_float2(w, w);
#pragma unroll
              for (int q = 0; q < Q; ++q) {
                const float4 a = lds128(p ^ ((uint32_t)q << 4));
                a2[2 * q] = __ffma2_rn(ww, make_float2(a.x, a.y), a2[2 * q]);
                a2[2 * q + 1] = __ffma2_rn(ww, make_float2(a.z, a.w), a2[2 * q + 1]);
              }
            }
          }
          if (k >= 2) {     // T_k = 2 L T_{k-1} - T_{k-2}; T_{k-2}'s own row is where T_k goes
            const float2 two = make_float2(2.0f, 2.0f);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
              const float4 o = lds128(((k & 1) ? own_o[q] : own_e[q]) + 32 * m * ROWB);
              a2[2 * q] = __ffma2_rn(two, a2[2 * q], make_float2(-o.x, -o.y));
              a2[2 * q + 1] = __ffma2_rn(two, a2[2 * q + 1], make_float2(-o.z, -o.w));
            }
          }
#pragma unroll
          for (int q = 0; q < Q; ++q)
            sts128(((k & 1) ? own_o[q] : own_e[q]) + 32 * m * ROWB,
                   make_float4(a2[2 * q].x, a2[2 * q].y, a2[2 * q + 1].x, a2[2 * q + 1].y));
        }
      }
      __syncwarp();
      if (k == 1) {   // order 0 is applied here: the Theta copy had the previous epilogue + one propagation to land
        mbar_wait_parity(a_bar + 16, (uint32_t)(it & 1));
        if constexpr (TRANS) mma_order_t<F, MTMAX>(acc, a_even + afr_t, bfr, MT);
        else mma_order<F, MTMAX>(acc, ta_e, bfr, MT);
      }
      if constexpr (TRANS) mma_order_t<F, MTMAX>(acc, ((k & 1) ? a_odd : a_even) + afr_t, bfr + (uint32_t)k * F * F * 4, MT);
      else mma_order<F, MTMAX>(acc, (k & 1) ? ta_o : ta_e, bfr + (uint32_t)k * F * F * 4, MT);
    }
    if (K == 1) {
      mbar_wait_parity(a_bar + 16, (uint32_t)(it & 1));
      if constexpr (TRANS) mma_order_t<F, MTMAX>(acc, a_even + afr_t, bfr, MT);
      else mma_order<F, MTMAX>(acc, ta_e, bfr, MT);
    }
    __syncwarp();            // every lane has read its last Theta / T fragment
    issue_theta(it + 1);
    // ---- epilogue: + bias, dense slab in the odd buffer, one bulk store (rows >= n of a tile are never stored)
#pragma unroll
    for (int mt = 0; mt < MTMAX; ++mt) {
      if (mt < MT) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const uint32_t o = ofr + ((16 * mt) * F + 8 * nt) * 4;
          sts64(o, make_float2(acc[mt][nt][0] + bias_m[nt][0], acc[mt][nt][1] + bias_m[nt][1]));
          sts64(o + 8 * ROWB, make_float2(acc[mt][nt][2] + bias_m[nt][0], acc[mt][nt][3] + bias_m[nt][1]));
        }
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0 && n > 0) {
      bulk_s2g(out + (size_t)d0.r0 * F, a_odd, (uint32_t)n * ROWB);
      bulk_commit();
    }
    d0 = d1;
    d1 = d2;
    d2 = d3;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
}

template <int F, int RPL, int K, bool TRANS>
static int launch(const Cfg& c, const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                  const int32_t* graph_ptr, const float* theta, int64_t sk, int64_t sg, const float* bias, float* out,
                  int64_t R, int64_t G, int32_t* meta, int max_nodes, cudaStream_t st) {
  auto kern = cheb_fwd_lane_kernel<F, RPL, K, TRANS>;
  FETA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
  int64_t grid = ceil_div(G, c.warps);
  if (grid > kNumSMs) grid = kNumSMs;
  kern<<<(unsigned)grid, c.warps * 32, c.smem, st>>>(x, rowptr, colidx, vals, graph_ptr, theta, sk, sg, bias, out, R, G,
                                                    c.nnz_cap, c.rows_cap, c.per_warp, meta, max_nodes);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}


// =====================================================================================================
// dTheta_k[g] = T_k^T . dOut over the rows of graph g (T_k recomputed by the forward recursion): per order one
// [F x n] . [n x F] product on the tensor cores (3xTF32), M = filter row i, N = filter column j, contraction over
// the graph's rows in k-steps of 8 (rows past the graph are zeroed once per graph in every slab).  F = 16: the
// channel <-> fragment-index maps are chosen so that every fragment is a 64-bit load (m = g, g + 8 <-> channels
// 2g, 2g + 1; n-tile nt, column n' <-> channel 2n' + nt) and a lane ends up with four consecutive j of one filter
// row -- the [K, F, F] block is assembled in shared memory and leaves with one bulk store.
// =====================================================================================================
static Cfg config_dtheta(int F, int K, int max_nodes) {
  Cfg c{0, 0, 0, 0, 0, 0, false};
  if (!(F == 8 || F == 16) || K < 1 || K > 4 || max_nodes < 1 || max_nodes > 64) return c;
  c.rpl = max_nodes <= 32 ? 1 : 2;
  c.rows_cap = (max_nodes + 15) / 16 * 16;
  c.nnz_cap = c.rows_cap * 4;
  const size_t slab = (size_t)c.rows_cap * F * 4;
  const size_t stage = 2 * slab + 2 * csr_bytes_of(c.nnz_cap);
  c.per_warp = (uint32_t)align_up(kBarBytes + slab + (size_t)K * F * F * 4 + 2 * stage, 128);
  int w = (int)((227 * 1024) / c.per_warp);
  if (w > 12) w = 12;
  if (w < 4) return c;
  c.warps = w;
  c.smem = (size_t)c.per_warp * w;
  c.ok = true;
  return c;
}

template <int F, int RPL, int K>
__global__ void __launch_bounds__(384, 1) cheb_dtheta_lane_kernel(
    const float* __restrict__ x, const float* __restrict__ dout, const int32_t* __restrict__ rowptr,
    const int32_t* __restrict__ colidx, const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr,
    float* __restrict__ dtheta, int64_t sk, int64_t sg, int64_t R, int64_t G, int nnz_cap, int rows_cap,
    uint32_t per_warp_bytes, int32_t* meta, int max_nodes) {
  constexpr int Q = F / 4, NT = F / 8, KSMAX = 4 * RPL;
  constexpr uint32_t ROWB = F * 4;
  constexpr bool CACHE_B = RPL == 1;        // dOut fragments stay in registers across the K orders
  extern __shared__ __align__(128) unsigned char sm[];
  if (!plan_guard_ok(meta, G, max_nodes)) { nan_fill_theta(dtheta, sk, sg, K, G, F * F); return; }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * warps + warp, stride = (int64_t)gridDim.x * warps;
  if (gw >= G) return;
  const int n_it = (int)((G - gw + stride - 1) / stride);

  const uint32_t slab = (uint32_t)rows_cap * ROWB;
  const uint32_t csr_bytes = (uint32_t)(((nnz_cap + 8) * 4 + 63) / 64 * 64);
  constexpr uint32_t th_bytes = (uint32_t)K * F * F * 4;
  const uint32_t base = smem_u32(sm) + (uint32_t)warp * per_warp_bytes;
  const uint32_t a_bar = base, a_odd = base + kBarBytes, a_out = a_odd + slab, a_st0 = a_out + th_bytes;
  const uint32_t stage_bytes = 2 * slab + 2 * csr_bytes;
  const bool out_contig = (sk == (int64_t)F * F);

  if (lane == 0) {
    mbar_init(a_bar, 1);
    mbar_init(a_bar + 8, 1);
    mbar_fence_init();
  }
  __syncwarp();
  const int nnz_total = __ldg(rowptr + R);

  const uint32_t g = (uint32_t)lane >> 2, tq = (uint32_t)lane & 3u;
  const uint32_t sw_row = swz<F>((uint32_t)lane);
  uint32_t own[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) own[q] = (uint32_t)lane * ROWB + (((uint32_t)q ^ sw_row) << 4);
  // fragment offsets of contraction rows tq (h = 0) and tq + 4 (h = 1); k-step ks adds 8 rows (same swizzle)
  uint32_t afr[2], bfr[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t r = tq + 4u * h;
    if (F == 16) {
      afr[h] = r * ROWB + (((g >> 1) ^ swz<F>(r)) << 4) + (g & 1u) * 8;   // T_k[r][2g, 2g + 1]
      bfr[h] = r * ROWB + g * 8;                                           // dOut[r][2g, 2g + 1] (dense slab)
    } else {
      afr[h] = r * ROWB + (((g >> 2) ^ swz<F>(r)) << 4) + (g & 3u) * 4;   // T_k[r][g]
      bfr[h] = r * ROWB + g * 4;                                           // dOut[r][g]
    }
  }
  const uint32_t ofr = F == 16 ? a_out + (2 * g * F + 4 * tq) * 4 : a_out + (g * F + 2 * tq) * 4;

  auto fetch_a = [&](int it, GraphDesc<RPL>& d) {
    d.r0 = d.r1 = 0;
    if (it < n_it) {
      const int64_t gi = gw + (int64_t)it * stride;
      d.r0 = __ldg(graph_ptr + gi);
      d.r1 = __ldg(graph_ptr + gi + 1);
    }
  };
  auto fetch_b = [&](int it, GraphDesc<RPL>& d) {
    d.e_lo = d.e_hi = 0;
#pragma unroll
    for (int m = 0; m < RPL; ++m) d.e0[m] = d.e1[m] = 0;
    if (it < n_it) {
      d.e_lo = __ldg(rowptr + d.r0);
      d.e_hi = __ldg(rowptr + d.r1);
#pragma unroll
      for (int m = 0; m < RPL; ++m) {
        const int row = lane + 32 * m;
        if (row < d.r1 - d.r0) {
          d.e0[m] = __ldg(rowptr + d.r0 + row);
          d.e1[m] = __ldg(rowptr + d.r0 + row + 1);
        }
      }
    }
  };
  auto staged_csr = [&](const GraphDesc<RPL>& d, int& a_lo, int& a_hi) {
    a_lo = d.e_lo & ~3;
    a_hi = (d.e_hi + 3) & ~3;
    return (a_hi - a_lo <= nnz_cap) && (a_hi <= nnz_total);
  };
  auto issue = [&](int it, const GraphDesc<RPL>& d) {
    if (it >= n_it) return;
    const uint32_t st = a_st0 + (uint32_t)(it & 1) * stage_bytes;
    const uint32_t bar = a_bar + 8u * (uint32_t)(it & 1);
    const int n = d.r1 - d.r0;
    int a_lo, a_hi;
    const bool copy_csr = staged_csr(d, a_lo, a_hi) && (d.e_hi > d.e_lo);
    const uint32_t bytes = 2u * (uint32_t)n * ROWB + (copy_csr ? 2u * (uint32_t)(a_hi - a_lo) * 4u : 0u);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive_expect_tx(bar, bytes);
      if (n > 0) {
        bulk_g2s(st, x + (size_t)d.r0 * F, (uint32_t)n * ROWB, bar);
        bulk_g2s(st + slab, dout + (size_t)d.r0 * F, (uint32_t)n * ROWB, bar);
      }
      if (copy_csr) {
        bulk_g2s(st + 2 * slab, colidx + a_lo, (uint32_t)(a_hi - a_lo) * 4, bar);
        bulk_g2s(st + 2 * slab + csr_bytes, vals + a_lo, (uint32_t)(a_hi - a_lo) * 4, bar);
      }
    }
  };

  GraphDesc<RPL> d0, d1, d2, d3;
  fetch_a(0, d0);
  fetch_a(1, d1);
  fetch_a(2, d2);
  fetch_b(0, d0);
  fetch_b(1, d1);
  issue(0, d0);

  for (int it = 0; it < n_it; ++it) {
    fetch_a(it + 3, d3);
    fetch_b(it + 2, d2);
    issue(it + 1, d1);

    const uint32_t a_even = a_st0 + (uint32_t)(it & 1) * stage_bytes;
    const uint32_t a_d = a_even + slab, a_ci = a_d + slab, a_cv = a_ci + csr_bytes;
    const int n = d0.r1 - d0.r0;
    const int KSu = (n + 7) >> 3;
    int a_lo, a_hi;
    const bool staged = staged_csr(d0, a_lo, a_hi);
    uint32_t own_e[Q], own_o[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) own_e[q] = a_even + own[q], own_o[q] = a_odd + own[q];
    mbar_wait_parity(a_bar + 8u * (uint32_t)(it & 1), (uint32_t)((it >> 1) & 1));
    if (lane == 0) bulk_wait_read0();   // the previous graph's dTheta store has drained the staging block
    __syncwarp();

    uint32_t nb[RPL][kSlots];
    float wt[RPL][kSlots];
    int deg[RPL], nslot[RPL];
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
      deg[m] = d0.e1[m] - d0.e0[m];
      nslot[m] = staged ? min(deg[m], kSlots) : 0;
      const uint32_t el = (uint32_t)(d0.e0[m] - a_lo) * 4;
#pragma unroll
      for (int j = 0; j < kSlots; ++j) {
        nb[m][j] = 0;
        wt[m][j] = 0.0f;
      }
      if (nslot[m] > 0) {
        const uint32_t c0 = (uint32_t)(lds32i(a_ci + el + 0) - d0.r0), c1 = (uint32_t)(lds32i(a_ci + el + 4) - d0.r0);
        const uint32_t c2 = (uint32_t)(lds32i(a_ci + el + 8) - d0.r0), c3 = (uint32_t)(lds32i(a_ci + el + 12) - d0.r0);
        wt[m][0] = lds32(a_cv + el + 0), wt[m][1] = lds32(a_cv + el + 4);
        wt[m][2] = lds32(a_cv + el + 8), wt[m][3] = lds32(a_cv + el + 12);
        nb[m][0] = c0 * ROWB + (swz<F>(c0) << 4), nb[m][1] = c1 * ROWB + (swz<F>(c1) << 4);
        nb[m][2] = c2 * ROWB + (swz<F>(c2) << 4), nb[m][3] = c3 * ROWB + (swz<F>(c3) << 4);
      }
    }
    // T_0: re-lay the own row of x (swizzled, in place); rows n .. 8 KSu - 1 enter the products: zero them everywhere
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
      const int row = lane + 32 * m;
      if (row < n) {
        float4 v[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) v[q] = lds128(a_even + (uint32_t)row * ROWB + 16 * q);
#pragma unroll
        for (int q = 0; q < Q; ++q) sts128(own_e[q] + 32 * m * ROWB, v[q]);
      } else if (row < 8 * KSu) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          sts128(own_e[q] + 32 * m * ROWB, z);
          sts128(own_o[q] + 32 * m * ROWB, z);
          sts128(a_d + (uint32_t)row * ROWB + 16 * q, z);
        }
      }
    }
    __syncwarp();

    uint32_t bh[CACHE_B ? KSMAX : 1][NT][2], bl[CACHE_B ? KSMAX : 1][NT][2];
    auto load_b = [&](int ks, uint32_t (&h_)[NT][2], uint32_t (&l_)[NT][2]) {
      if constexpr (F == 16) {
        const float2 w0 = lds64(a_d + bfr[0] + ks * 8 * ROWB), w1 = lds64(a_d + bfr[1] + ks * 8 * ROWB);
        split_tf32(w0.x, h_[0][0], l_[0][0]);
        split_tf32(w1.x, h_[0][1], l_[0][1]);
        split_tf32(w0.y, h_[1][0], l_[1][0]);
        split_tf32(w1.y, h_[1][1], l_[1][1]);
      } else {
        split_tf32(lds32(a_d + bfr[0] + ks * 8 * ROWB), h_[0][0], l_[0][0]);
        split_tf32(lds32(a_d + bfr[1] + ks * 8 * ROWB), h_[0][1], l_[0][1]);
      }
    };
    if constexpr (CACHE_B) {
#pragma unroll
      for (int ks = 0; ks < KSMAX; ++ks) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) bh[ks][nt][0] = bh[ks][nt][1] = bl[ks][nt][0] = bl[ks][nt][1] = 0u;
        if (ks < KSu) load_b(ks, bh[ks], bl[ks]);
      }
    }

#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (k >= 1) {
        const uint32_t a_src = (k & 1) ? a_even : a_odd;
#pragma unroll
        for (int m = 0; m < RPL; ++m) {
          if (lane + 32 * m < n) {
            float2 a2[F / 2];
#pragma unroll
            for (int i = 0; i < F / 2; ++i) a2[i] = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < kSlots; ++j) {
              if (j < nslot[m]) {
                const uint32_t p = a_src + nb[m][j];
                const float2 ww = make_float2(wt[m][j], wt[m][j]);
                float4 a[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) a[q] = lds128(p ^ ((uint32_t)q << 4));
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                  a2[2 * q] = __ffma2_rn(ww, make_float2(a[q].x, a[q].y), a2[2 * q]);
                  a2[2 * q + 1] = __ffma2_rn(ww, make_float2(a[q].z, a[q].w), a2[2 * q + 1]);
                }
              }
            }
            if (deg[m] > nslot[m]) {
              for (int e = d0.e0[m] + nslot[m]; e < d0.e1[m]; ++e) {
                uint32_t c;
                float w;
                if (staged) {
                  c = (uint32_t)(lds32i(a_ci + (uint32_t)(e - a_lo) * 4) - d0.r0);
                  w = lds32(a_cv + (uint32_t)(e - a_lo) * 4);
                } else {
                  c = (uint32_t)(__ldg(colidx + e) - d0.r0);
                  w = __ldg(vals + e);
                }
                const uint32_t p = a_src + c * ROWB + (swz<F>(c) << 4);
                const float2 ww = make_float2(w, w);
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                  const float4 a = lds128(p ^ ((uint32_t)q << 4));
                  a2[2 * q] = __ffma2_rn(ww, make_float2(a.x, a.y), a2[2 * q]);
                  a2[2 * q + 1] = __ffma2_rn(ww, make_float2(a.z, a.w), a2[2 * q + 1]);
                }
              }
            }
            if (k >= 2) {
              const float2 two = make_float2(2.0f, 2.0f);
#pragma unroll
              for (int q = 0; q < Q; ++q) {
                const float4 o = lds128(((k & 1) ? own_o[q] : own_e[q]) + 32 * m * ROWB);
                a2[2 * q] = __ffma2_rn(two, a2[2 * q], make_float2(-o.x, -o.y));
                a2[2 * q + 1] = __ffma2_rn(two, a2[2 * q + 1], make_float2(-o.z, -o.w));
              }
            }
#pragma unroll
            for (int q = 0; q < Q; ++q)
              sts128(((k & 1) ? own_o[q] : own_e[q]) + 32 * m * ROWB,
                     make_float4(a2[2 * q].x, a2[2 * q].y, a2[2 * q + 1].x, a2[2 * q + 1].y));
          }
        }
        __syncwarp();
      }
      // ---- dTheta_k = T_k^T dOut
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[nt][q] = 0.0f;
      const uint32_t a_t = (k & 1) ? a_odd : a_even;
#pragma unroll
      for (int ks = 0; ks < KSMAX; ++ks) {
        if (ks < KSu) {
          uint32_t ah[4], al[4];
          if constexpr (F == 16) {
            const float2 v0 = lds64(a_t + afr[0] + ks * 8 * ROWB), v1 = lds64(a_t + afr[1] + ks * 8 * ROWB);
            split_tf32(v0.x, ah[0], al[0]);
            split_tf32(v0.y, ah[1], al[1]);
            split_tf32(v1.x, ah[2], al[2]);
            split_tf32(v1.y, ah[3], al[3]);
          } else {
            split_tf32(lds32(a_t + afr[0] + ks * 8 * ROWB), ah[0], al[0]);
            split_tf32(lds32(a_t + afr[1] + ks * 8 * ROWB), ah[2], al[2]);
            ah[1] = ah[3] = al[1] = al[3] = 0u;     // M = 16 tile, filter rows 8 .. 15 do not exist
          }
          uint32_t th_[NT][2], tl_[NT][2];
          if constexpr (!CACHE_B) load_b(ks, th_, tl_);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], ah, CACHE_B ? bh[CACHE_B ? ks : 0][nt] : th_[nt]);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], ah, CACHE_B ? bl[CACHE_B ? ks : 0][nt] : tl_[nt]);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], al, CACHE_B ? bh[CACHE_B ? ks : 0][nt] : th_[nt]);
        }
      }
      if constexpr (F == 16) {
        sts128(ofr + (uint32_t)k * F * F * 4, make_float4(acc[0][0], acc[1][0], acc[0][1], acc[1][1]));
        sts128(ofr + (uint32_t)k * F * F * 4 + ROWB, make_float4(acc[0][2], acc[1][2], acc[0][3], acc[1][3]));
      } else {
        sts64(ofr + (uint32_t)k * F * F * 4, make_float2(acc[0][0], acc[0][1]));
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      const int64_t gi = gw + (int64_t)it * stride;
      if (out_contig) {
        bulk_s2g(dtheta + gi * sg, a_out, th_bytes);
      } else {
        for (int k = 0; k < K; ++k) bulk_s2g(dtheta + gi * sg + (int64_t)k * sk, a_out + (uint32_t)k * F * F * 4, F * F * 4);
      }
      bulk_commit();
    }
    d0 = d1;
    d1 = d2;
    d2 = d3;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int F, int RPL, int K>
static int launch_dtheta(const Cfg& c, const float* x, const float* dout, const int32_t* rowptr, const int32_t* colidx,
                         const float* vals, const int32_t* graph_ptr, float* dtheta, int64_t sk, int64_t sg, int64_t R,
                         int64_t G, int32_t* meta, int max_nodes, cudaStream_t st) {
  auto kern = cheb_dtheta_lane_kernel<F, RPL, K>;
  FETA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
  int64_t grid = ceil_div(G, c.warps);
  if (grid > kNumSMs) grid = kNumSMs;
  kern<<<(unsigned)grid, c.warps * 32, c.smem, st>>>(x, dout, rowptr, colidx, vals, graph_ptr, dtheta, sk, sg, R, G,
                                                    c.nnz_cap, c.rows_cap, c.per_warp, meta, max_nodes);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

}  // namespace lane

static int lane_dispatch(bool trans, const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                         const int32_t* graph_ptr, const float* theta, int64_t sk, int64_t sg, const float* bias,
                         float* out, int64_t R, int64_t G, int K, int F, int max_nodes, int32_t* meta, cudaStream_t st) {
  if (getenv("FETA_CHEB_NO_LANE_KERNEL") != nullptr) return 1;
  lane::Cfg c = lane::config(F, K, max_nodes);
  if (!c.ok) return 1;
  if (((uintptr_t)colidx % 16) || ((uintptr_t)vals % 16)) return 1;
#define FETA_LANE_CASE(F_, R_, K_)                                                                                    \
  if (F == F_ && c.rpl == R_ && K == K_) {                                                                            \
    if (trans)                                                                                                        \
      return lane::launch<F_, R_, K_, true>(c, x, rowptr, colidx, vals, graph_ptr, theta, sk, sg, bias, out, R, G,    \
                                            meta, max_nodes, st);                                                     \
    return lane::launch<F_, R_, K_, false>(c, x, rowptr, colidx, vals, graph_ptr, theta, sk, sg, bias, out, R, G,     \
                                           meta, max_nodes, st);                                                      \
  }
#define FETA_LANE_CASES(F_, R_) FETA_LANE_CASE(F_, R_, 1) FETA_LANE_CASE(F_, R_, 2) FETA_LANE_CASE(F_, R_, 3) FETA_LANE_CASE(F_, R_, 4)
  FETA_LANE_CASES(8, 1) FETA_LANE_CASES(8, 2) FETA_LANE_CASES(16, 1) FETA_LANE_CASES(16, 2)
#undef FETA_LANE_CASES
#undef FETA_LANE_CASE
  return 1;
}

// returns FETA_OK if launched, 1 if this shape is not eligible (the caller falls back), < 0 on error
int cheb_fwd_lane_try(const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                      const int32_t* graph_ptr, const float* theta, int64_t sk, int64_t sg, const float* bias,
                      float* out, int64_t R, int64_t G, int K, int F, int max_nodes, int32_t* meta, cudaStream_t st) {
  return lane_dispatch(false, x, rowptr, colidx, vals, graph_ptr, theta, sk, sg, bias, out, R, G, K, F, max_nodes, meta,
                       st);
}

// dx = sum_k T_k(L^T) dOut . Theta_k^T: the forward recursion over the SOURCE-grouped CSR against transposed filters
int cheb_bwd_dx_lane_try(const float* dout, const int32_t* rowptr_t, const int32_t* colidx_t, const float* vals_t,
                         const int32_t* graph_ptr, const float* theta, int64_t sk, int64_t sg, float* dx, int64_t R,
                         int64_t G, int K, int F, int max_nodes, int32_t* meta, cudaStream_t st) {
  return lane_dispatch(true, dout, rowptr_t, colidx_t, vals_t, graph_ptr, theta, sk, sg, nullptr, dx, R, G, K, F,
                       max_nodes, meta, st);
}

// dTheta_k[g] = T_k^T dOut over the rows of graph g (T_k recomputed)
int cheb_bwd_dtheta_lane_try(const float* x, const float* dout, const int32_t* rowptr, const int32_t* colidx,
                             const float* vals, const int32_t* graph_ptr, float* dtheta, int64_t sk, int64_t sg,
                             int64_t R, int64_t G, int K, int F, int max_nodes, int32_t* meta, cudaStream_t st) {
  if (getenv("FETA_CHEB_NO_LANE_KERNEL") != nullptr) return 1;
  lane::Cfg c = lane::config_dtheta(F, K, max_nodes);
  if (!c.ok) return 1;
  if (((uintptr_t)colidx % 16) || ((uintptr_t)vals % 16) || sk != (int64_t)F * F) return 1;
#define FETA_LANE_CASE(F_, R_, K_)                                                                                    \
  if (F == F_ && c.rpl == R_ && K == K_)                                                                              \
    return lane::launch_dtheta<F_, R_, K_>(c, x, dout, rowptr, colidx, vals, graph_ptr, dtheta, sk, sg, R, G, meta,   \
                                           max_nodes, st);
#define FETA_LANE_CASES(F_, R_) FETA_LANE_CASE(F_, R_, 1) FETA_LANE_CASE(F_, R_, 2) FETA_LANE_CASE(F_, R_, 3) FETA_LANE_CASE(F_, R_, 4)
  FETA_LANE_CASES(8, 1) FETA_LANE_CASES(8, 2) FETA_LANE_CASES(16, 1) FETA_LANE_CASES(16, 2)
#undef FETA_LANE_CASES
#undef FETA_LANE_CASE
  return 1;
}

}  // namespace feta
