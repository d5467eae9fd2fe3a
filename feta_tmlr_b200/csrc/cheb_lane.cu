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
// lanes past a graph's last row still run the (branch-free) gather on their own row's address with weight 0; that
// address may lie up to 64 rows past a slab -- inside the next warp's region, or, for the last warp, this pad
constexpr size_t kTailPad = 4096;

static inline uint32_t csr_bytes_of(int nnz_cap) { return (uint32_t)align_up((size_t)(nnz_cap + 8) * 4, 64); }

static Cfg config(int F, int K, int max_nodes) {
  Cfg c{0, 0, 0, 0, 0, 0, false};
  if (!(F == 8 || F == 16) || K < 1 || K > 4 || max_nodes < 1 || max_nodes > 64) return c;
  c.rpl = max_nodes <= 32 ? 1 : 2;
  // slabs hold whole 8-row groups; the last m16 tile of a graph may READ up to 8 rows past its slab (the next
  // region of the same warp; those output rows are never stored) but nothing is written there
  c.rows_cap = (max_nodes + 7) / 8 * 8;
  c.nnz_cap = c.rows_cap * 3;                     // molecules: ~2.2 entries per row; larger slices are read through L2
  if (const char* e = getenv("FETA_LANE_NNZ_PER_ROW")) c.nnz_cap = c.rows_cap * atoi(e);
  const size_t slab = (size_t)c.rows_cap * F * 4;
  const size_t stage = slab + 2 * csr_bytes_of(c.nnz_cap);
  c.per_warp = (uint32_t)align_up(kBarBytes + slab + (size_t)K * F * F * 4 + 2 * stage, 128);
  int w = (int)((227 * 1024 - kTailPad) / c.per_warp);
  int wmax = 16;
  if (const char* e = getenv("FETA_LANE_WARPS")) wmax = atoi(e);
  if (w > wmax) w = wmax;
  if (w < 4) return c;
  c.warps = w;
  c.smem = (size_t)c.per_warp * w + kTailPad;
  c.ok = true;
  return c;
}

template <int RPL>
struct GraphDesc {  // scalars of one graph, fetched ahead of use
  int r0, r1, e_lo, e_hi;
  int e0[RPL], e1[RPL];
};

// acc[mt][nt] += T[16mt .. 16mt+15, :] . Theta_k (or Theta_k^T), 3xTF32: hi.hi + hi.lo + lo.hi.
// Forward variant, scalar fragment loads in the textbook mapping (k-step ks, half h -> contraction index
// 8 ks + 4 h + tq).  `ta[j]` = address of this lane's word in 16-byte chunk j of row g (rows 16 mt + g, + 8 are
// immediates); `tb` = address of Theta_k[tq][first own output channel].  F = 16 permutes the OUTPUT channels (n-tile
// nt, column n' <-> channel 2 n' + nt) so that a lane ends up with four consecutive channels of a row.
// Vector loads are NOT a win here: HMMA wants {a0..a3} in one aligned register quad and {b0, b1} in a pair, and
// values that arrive through one LDS.128 / LDS.64 per row have to be MOVed into place (measured: -100 LDS,
// +270 MOV/IMAD per graph) -- except where one load delivers the pair in order, as for Theta^T below.
template <int F, int MTMAX>
__device__ __forceinline__ void mma_order(float (&acc)[MTMAX][F / 8][4], const uint32_t (&ta)[F / 4], uint32_t tb,
                                          int MT) {
  constexpr int NT = F / 8, KS = F / 8;
  uint32_t bh[NT][KS][2], bl[NT][KS][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      split_tf32(lds32(tb + ((8 * ks) * F + nt) * 4), bh[nt][ks][0], bl[nt][ks][0]);
      split_tf32(lds32(tb + ((8 * ks + 4) * F + nt) * 4), bh[nt][ks][1], bl[nt][ks][1]);
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

// The product against Theta_k^T (the dx kernel) with vector fragment loads: the contraction index of an MMA may be
// permuted freely as long as A and B agree; here lane (g, tq) owns the CONTIGUOUS indices F/4 * tq .. F/4 * tq +
// F/4 - 1 (k-step ks, half h -> F/4 * tq + 2 ks + h), so its B values are one vector load of a filter row that
// delivers every {b0, b1} pair in register order, conflict-free (a scalar B-fragment load of Theta^T would be 8-way
// bank conflicted); its A values are scalar loads (a vector load would have to be MOVed into the {a0..a3} quad).
// `ta` = address of (row g, first own channel) in the slab, `tb` = address of the filter row's first own channel.
template <int F, int MTMAX>
__device__ __forceinline__ void mma_order_t(float (&acc)[MTMAX][F / 8][4], uint32_t ta, uint32_t tb, int MT) {
  constexpr int NT = F / 8, KS = F / 8, V = F / 4;   // V own contraction indices per lane
  uint32_t bh[NT][KS][2], bl[NT][KS][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    float v[V];
    if constexpr (F == 16) {
      const float4 t = lds128(tb + (2 * nt) * F * 4);   // filter row 4 (g / 2) + 2 nt + g % 2, see the kernel
      v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    } else {
      const float2 t = lds64(tb);
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
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        // scalar loads straight into the {a0..a3} register quad (own indices 2 ks, 2 ks + 1 of rows g, g + 8)
        uint32_t ah[4], al[4];
        split_tf32(lds32(ta + (16 * mt) * F * 4 + (2 * ks) * 4), ah[0], al[0]);
        split_tf32(lds32(ta + (16 * mt + 8) * F * 4 + (2 * ks) * 4), ah[1], al[1]);
        split_tf32(lds32(ta + (16 * mt) * F * 4 + (2 * ks + 1) * 4), ah[2], al[2]);
        split_tf32(lds32(ta + (16 * mt + 8) * F * 4 + (2 * ks + 1) * 4), ah[3], al[3]);
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

// The edges of one own row, decoded once per graph: shared-memory offset of the neighbour row (chunk 0, swizzled)
// inside a slab + weight.
struct RowEdges {
  uint32_t nb[kSlots];
  float wt[kSlots];
  int e0, e1, nslot;
};

template <int F>
__device__ __forceinline__ void decode_edges(RowEdges& re, int e0, int e1, bool staged, uint32_t a_ci, uint32_t a_cv,
                                             int a_lo, int r0) {
  constexpr uint32_t ROWB = F * 4;
  re.e0 = e0;
  re.e1 = e1;
  re.nslot = staged ? min(e1 - e0, kSlots) : 0;         // e1 == e0 for lanes past the graph
#pragma unroll
  for (int j = 0; j < kSlots; ++j) re.nb[j] = 0u, re.wt[j] = 0.0f;
  if (re.nslot > 0) {   // up to 3 entries past the row are read (inside the staged slice + slack), never used
    const uint32_t el = (uint32_t)(e0 - a_lo) * 4;
#pragma unroll
    for (int j = 0; j < kSlots; ++j) {
      const uint32_t c = (uint32_t)(lds32i(a_ci + el + 4 * j) - r0);
      re.wt[j] = lds32(a_cv + el + 4 * j);
      re.nb[j] = c * ROWB + (swz<F>(c) << 4);
    }
  }
}

template <int F>
__device__ __forceinline__ void gather_slot(float2 (&a2)[F / 2], uint32_t p, float w, bool first) {
  constexpr int Q = F / 4;
  const float2 ww = make_float2(w, w);
  float4 a[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) a[q] = lds128(p ^ ((uint32_t)q << 4));   // 64-byte aligned slabs: the XOR stays in the row
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    a2[2 * q] = __ffma2_rn(ww, make_float2(a[q].x, a[q].y), first ? make_float2(0.f, 0.f) : a2[2 * q]);
    a2[2 * q + 1] = __ffma2_rn(ww, make_float2(a[q].z, a[q].w), first ? make_float2(0.f, 0.f) : a2[2 * q + 1]);
  }
}

// One row of T_k = c L T_{k-1} (- T_{k-2}): the decoded slots, then longer rows (or a CSR slice that was not staged)
// in a generic loop; `dst` = the own row's chunk addresses in the buffer that holds T_{k-2} and receives T_k.
template <int F, bool SUB>
__device__ __forceinline__ void propagate_row(const RowEdges& re, uint32_t a_src, const uint32_t* dst,
                                              uint32_t dst_off, bool staged,
                                              uint32_t a_ci, uint32_t a_cv, int a_lo, int r0,
                                              const int32_t* __restrict__ colidx, const float* __restrict__ vals) {
  constexpr int Q = F / 4;
  constexpr uint32_t ROWB = F * 4;
  float2 a2[F / 2];
#pragma unroll
  for (int i = 0; i < F / 2; ++i) a2[i] = make_float2(0.f, 0.f);
  // per-slot divergent branches on purpose (a branch-free variant with weight-0 dummy slots measured 23 % slower:
  // more instructions for the many lanes / slots without an edge)
#pragma unroll
  for (int j = 0; j < kSlots; ++j)
    if (j < re.nslot) gather_slot<F>(a2, a_src + re.nb[j], re.wt[j], false);
  if (re.e1 - re.e0 > re.nslot) {
    for (int e = re.e0 + re.nslot; e < re.e1; ++e) {
      uint32_t c;
      float w;
      if (staged) {
        c = (uint32_t)(lds32i(a_ci + (uint32_t)(e - a_lo) * 4) - r0);
        w = lds32(a_cv + (uint32_t)(e - a_lo) * 4);
      } else {
        c = (uint32_t)(__ldg(colidx + e) - r0);
        w = __ldg(vals + e);
      }
      gather_slot<F>(a2, a_src + c * ROWB + (swz<F>(c) << 4), w, false);
    }
  }
  if (SUB) {     // T_k = 2 L T_{k-1} - T_{k-2}; T_{k-2}'s own row is where T_k goes
    const float2 two = make_float2(2.0f, 2.0f);
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const float4 o = lds128(dst[q] + dst_off);
      a2[2 * q] = __ffma2_rn(two, a2[2 * q], make_float2(-o.x, -o.y));
      a2[2 * q + 1] = __ffma2_rn(two, a2[2 * q + 1], make_float2(-o.z, -o.w));
    }
  }
#pragma unroll
  for (int q = 0; q < Q; ++q)
    sts128(dst[q] + dst_off, make_float4(a2[2 * q].x, a2[2 * q].y, a2[2 * q + 1].x, a2[2 * q + 1].y));
}

// The dense [n, F] slab a bulk copy landed -> XOR-swizzled rows, in place, each lane its own row(s): the permutation
// stays inside a row, so no synchronisation is needed.  (A coalesced lane-per-chunk pass would avoid the 4-way bank
// conflict of the dense read, but costs two __syncwarp and ~40 more instructions per graph -- the kernel is bound by
// the per-warp instruction rate, not by shared-memory wavefronts: measured slower.)
template <int F, int RPL>
__device__ __forceinline__ void relay_rows(uint32_t a_slab, int n, int lane, const uint32_t (&own)[F / 4]) {
  constexpr int Q = F / 4;
  constexpr uint32_t ROWB = F * 4;
#pragma unroll
  for (int m = 0; m < RPL; ++m) {
    if (lane + 32 * m < n) {
      float4 v[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) v[q] = lds128(a_slab + (uint32_t)(lane + 32 * m) * ROWB + 16 * q);
#pragma unroll
      for (int q = 0; q < Q; ++q) sts128(a_slab + own[q] + 32 * m * ROWB, v[q]);
    }
  }
  __syncwarp();
}

// TRANS = false: out = sum_k T_k(L) x . Theta_k + bias (forward).  TRANS = true: the same recursion against
// Theta_k^T -- called with (dOut, the source-grouped CSR = L^T, no bias) it is the input gradient, because the
// operator acts on rows and the filters on channels: dx = sum_k T_k(L^T) dOut . Theta_k^T.
template <int F, int RPL, int K, bool TRANS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) cheb_fwd_lane_kernel(
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
  // A fragments: one vector of own contraction indices of row g (see mma_order_t)
  const uint32_t afr_t = F == 16 ? g * ROWB + ((tq ^ swz<F>(g)) << 4) : g * ROWB + (((tq >> 1) ^ swz<F>(g)) << 4) + (tq & 1u) * 8;
  // B fragments.  F = 16 permutes the OUTPUT channels so that a lane ends up with four consecutive channels of a row
  // (one STS.128 per row in the epilogue instead of two 4-way conflicted STS.64): forward -- n-tile nt, column n'
  // <-> channel 2 n' + nt (Theta_k[tq][2g, 2g + 1] is one 64-bit load); TRANS -- <-> 4 (n' / 2) + 2 nt + n' % 2
  // (the quarter-warp's filter rows stay adjacent: conflict-free 128-bit loads).
  const uint32_t bfr = TRANS ? (F == 16 ? a_th + ((4 * (g >> 1) + (g & 1u)) * F + 4 * tq) * 4 : a_th + (g * F + 2 * tq) * 4)
                             : (F == 16 ? a_th + (tq * F + 2 * g) * 4 : a_th + (tq * F + g) * 4);
  const uint32_t ofr = F == 16 ? a_odd + (g * F + 4 * tq) * 4 : a_odd + (g * F + 2 * tq) * 4;   // output staging (dense)
  uint32_t afr[Q];                                               // forward A fragment: row g, chunk j, word tq
#pragma unroll
  for (int j = 0; j < Q; ++j) afr[j] = g * ROWB + (((uint32_t)j ^ swz<F>(g)) << 4) + tq * 4;
  uint32_t own_o[Q];                                             // the odd buffer never moves
#pragma unroll
  for (int q = 0; q < Q; ++q) own_o[q] = a_odd + own[q];
  float bias_m[F / 4];                                           // the F / 4 consecutive channels this lane stores
#pragma unroll
  for (int i = 0; i < F / 4; ++i) bias_m[i] = bias ? __ldg(bias + (F / 4) * tq + i) : 0.0f;

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
    uint32_t own_e[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) own_e[q] = a_even + own[q];
    mbar_wait_parity(a_bar + 8u * (uint32_t)(it & 1), (uint32_t)((it >> 1) & 1));
    if (lane == 0) bulk_wait_read0();   // the previous graph's output store has drained the odd buffer
    __syncwarp();

    // ---- decode the first kSlots CSR entries of every own row (neighbour-row offset + weight)
    RowEdges re[RPL];
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
      decode_edges<F>(re[m], d0.e0[m], d0.e1[m], staged, a_ci, a_cv, a_lo, d0.r0);
    }

    // ---- T_0: the dense x slab -> swizzled rows, in place
    relay_rows<F, RPL>(a_even, n, lane, own);

    float acc[MTMAX][NT][4];
#pragma unroll
    for (int mt = 0; mt < MTMAX; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.0f;

    auto mma_ord = [&](uint32_t a_buf, int k) {     // acc += T_k . Theta_k (or Theta_k^T) out of slab `a_buf`
      if constexpr (TRANS) {
        mma_order_t<F, MTMAX>(acc, a_buf + afr_t, bfr + (uint32_t)k * F * F * 4, MT);
      } else {
        uint32_t ta[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) ta[j] = a_buf + afr[j];
        mma_order<F, MTMAX>(acc, ta, bfr + (uint32_t)k * F * F * 4, MT);
      }
    };
#pragma unroll
    for (int k = 1; k < K; ++k) {
      const uint32_t a_src = (k & 1) ? a_even : a_odd;   // T_{k-1}
#pragma unroll
      for (int m = 0; m < RPL; ++m) {
        if (lane + 32 * m < n) {
          if (k >= 2)
            propagate_row<F, true>(re[m], a_src, (k & 1) ? own_o : own_e, 32 * m * ROWB, staged, a_ci, a_cv, a_lo, d0.r0, colidx, vals);
          else
            propagate_row<F, false>(re[m], a_src, (k & 1) ? own_o : own_e, 32 * m * ROWB, staged, a_ci, a_cv, a_lo, d0.r0, colidx, vals);
        }
      }
      __syncwarp();
      if (k == 1) {   // order 0 is applied here: the Theta copy had the previous epilogue + one propagation to land
        mbar_wait_parity(a_bar + 16, (uint32_t)(it & 1));
        mma_ord(a_even, 0);
      }
      mma_ord((k & 1) ? a_odd : a_even, k);
    }
    if (K == 1) {
      mbar_wait_parity(a_bar + 16, (uint32_t)(it & 1));
      mma_ord(a_even, 0);
    }
    __syncwarp();            // every lane has read its last Theta / T fragment
    issue_theta(it + 1);
    // ---- epilogue: + bias, dense slab in the odd buffer, one bulk store (rows >= n of a tile are never stored)
#pragma unroll
    for (int mt = 0; mt < MTMAX; ++mt) {
      if (mt < MT) {
        const uint32_t o = ofr + (16 * mt) * ROWB;
        const bool lo_ok = 16 * mt + (int)g < n, hi_ok = 16 * mt + 8 + (int)g < n;
        if constexpr (F == 16) {
          // channels 4 tq .. 4 tq + 3 of rows g, g + 8: forward (c0 nt0, c0 nt1, c1 nt0, c1 nt1); TRANS (c0 nt0, c1 nt0, c0 nt1, c1 nt1)
          const float (&a0)[4] = acc[mt][0];
          const float (&a1)[4] = acc[mt][NT - 1];
          if (lo_ok)
            sts128(o, TRANS ? make_float4(a0[0] + bias_m[0], a0[1] + bias_m[1], a1[0] + bias_m[2], a1[1] + bias_m[3])
                            : make_float4(a0[0] + bias_m[0], a1[0] + bias_m[1], a0[1] + bias_m[2], a1[1] + bias_m[3]));
          if (hi_ok)
            sts128(o + 8 * ROWB,
                   TRANS ? make_float4(a0[2] + bias_m[0], a0[3] + bias_m[1], a1[2] + bias_m[2], a1[3] + bias_m[3])
                         : make_float4(a0[2] + bias_m[0], a1[2] + bias_m[1], a0[3] + bias_m[2], a1[3] + bias_m[3]));
        } else {
          if (lo_ok) sts64(o, make_float2(acc[mt][0][0] + bias_m[0], acc[mt][0][1] + bias_m[1]));
          if (hi_ok) sts64(o + 8 * ROWB, make_float2(acc[mt][0][2] + bias_m[0], acc[mt][0][3] + bias_m[1]));
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

template <int F, int RPL, int K, bool TRANS, int MAXT = 512>
static int launch(const Cfg& c, const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                  const int32_t* graph_ptr, const float* theta, int64_t sk, int64_t sg, const float* bias, float* out,
                  int64_t R, int64_t G, int32_t* meta, int max_nodes, cudaStream_t st) {
  auto kern = cheb_fwd_lane_kernel<F, RPL, K, TRANS, MAXT>;
  FETA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
  const int warps = c.warps * 32 > MAXT ? MAXT / 32 : c.warps;
  int64_t grid = ceil_div(G, warps);
  if (grid > kNumSMs) grid = kNumSMs;
  kern<<<(unsigned)grid, warps * 32, (size_t)c.per_warp * warps + kTailPad, st>>>(x, rowptr, colidx, vals, graph_ptr, theta, sk, sg,
                                                                      bias, out, R, G, c.nnz_cap, c.rows_cap, c.per_warp,
                                                                      meta, max_nodes);
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
  c.rows_cap = (max_nodes + 7) / 8 * 8;         // the contraction walks whole 8-row k-steps
  c.nnz_cap = c.rows_cap * 3;
  if (const char* e = getenv("FETA_LANE_NNZ_PER_ROW")) c.nnz_cap = c.rows_cap * atoi(e);
  const size_t slab = (size_t)c.rows_cap * F * 4;
  const size_t stage = 2 * slab + 2 * csr_bytes_of(c.nnz_cap);
  c.per_warp = (uint32_t)align_up(kBarBytes + slab + (size_t)K * F * F * 4 + 2 * stage, 128);
  int w = (int)((227 * 1024 - kTailPad) / c.per_warp);
  int wmax = 12;
  if (const char* e = getenv("FETA_LANE_WARPS")) wmax = atoi(e) < 12 ? atoi(e) : 12;
  if (w > wmax) w = wmax;
  if (w < 4) return c;
  c.warps = w;
  c.smem = (size_t)c.per_warp * w + kTailPad;
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

    RowEdges re[RPL];
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
      decode_edges<F>(re[m], d0.e0[m], d0.e1[m], staged, a_ci, a_cv, a_lo, d0.r0);
    }
    // T_0: the dense x slab -> swizzled rows, in place; rows n .. 8 KSu - 1 enter the products: zero them in every slab
    relay_rows<F, RPL>(a_even, n, lane, own);
#pragma unroll
    for (int m = 0; m < RPL; ++m) {
      const int row = lane + 32 * m;
      if (row >= n && row < 8 * KSu) {
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
            if (k >= 2)
              propagate_row<F, true>(re[m], a_src, (k & 1) ? own_o : own_e, 32 * m * ROWB, staged, a_ci, a_cv, a_lo, d0.r0, colidx, vals);
            else
              propagate_row<F, false>(re[m], a_src, (k & 1) ? own_o : own_e, 32 * m * ROWB, staged, a_ci, a_cv, a_lo, d0.r0, colidx, vals);
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
      return lane::launch<F_, R_, K_, true, 512>(c, x, rowptr, colidx, vals, graph_ptr, theta, sk, sg, bias, out, R,  \
                                                 G, meta, max_nodes, st);                                             \
    /* measured (profiles/r2_cheb_lane.md): the F = 16 two-rows-per-lane forward is faster with 12 warps x 165   */  \
    /* registers than with 16 x 128; every other variant wants the 16 warps (13..15 cap registers at 128 too)    */  \
    return lane::launch<F_, R_, K_, false, (F_ == 16 && R_ == 2) ? 384 : 512>(c, x, rowptr, colidx, vals, graph_ptr, \
                                                                              theta, sk, sg, bias, out, R, G, meta,  \
                                                                              max_nodes, st);                        \
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
