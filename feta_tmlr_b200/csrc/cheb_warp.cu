// A1/A3: warp-per-graph, TMA-staged, persistent forward of ChebConvDynamic for batches of small
// graphs (every graph <= 64 rows): the HBM-roofline variant of csrc/cheb.cu.
//
// Why a second kernel: ncu on the chunk kernel (profiles/r1_cheb_fwd_ncu.md) showed 53 %
// long-scoreboard stalls (Theta / CSR read through L1 in the inner loop) and 38 % barrier stalls
// (idle threads + block-wide __syncthreads per Chebyshev order).  Here
//   * one WARP owns one graph at a time (lane = row, up to RPL rows per lane), so the only
//     synchronisation between Chebyshev orders is __syncwarp;
//   * everything a graph needs -- its x rows, its Theta block, its CSR slice -- is brought into
//     shared memory by cp.async.bulk (TMA, 1-D) completing on an mbarrier, double buffered, issued
//     one graph ahead; the scalars that size those copies (graph_ptr / rowptr reads) are fetched
//     two and three graphs ahead, so no global-load latency sits on the critical path;
//   * the filter is applied with packed fp32x2 FMAs (FFMA2) against Theta rows broadcast from
//     shared memory (F = 4, 8), or -- F = 16, where those broadcasts saturated the shared-memory
//     pipe -- as m16n8k8 TF32 MMAs with both operands split hi + lo (three MMAs per product,
//     fp32-grade accuracy) over XOR-swizzled 16-float slabs (conflict-free lane-per-row float4
//     accesses AND A-fragment loads); that variant keeps ONE Theta buffer with its own barrier and
//     propagates T_1 before the first filter application so the refill hides behind other work;
//   * the output slab goes back with one cp.async.bulk store per graph;
//   * the grid is persistent: 148 CTAs, each warp strides over the graph list.
// Measured history and the ncu captures behind each step: profiles/r1_cheb_fwd_ncu.md.
#include <stdlib.h>

#include "common.cuh"
#include "graph_tile.cuh"

namespace feta {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
// global -> shared bulk copy (TMA, UBLKCP), completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// shared -> global bulk store
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct WarpCfg {
  int rpl, warps, nnz_cap, rows_cap;
  bool mma;
  size_t per_warp, smem;
  bool ok;
};

static inline size_t warp_region_bytes(int F, int K, int rows_cap, int nnz_cap, bool mma) {
  // odd orders: padded rows (FFMA2 variant) or dense XOR-swizzled rows (MMA variant)
  const size_t ta = (size_t)rows_cap * (mma ? F : F + 4) * 4;
  const size_t t0 = (size_t)rows_cap * F * 4;        // x slab / even orders, dense (one bulk copy)
  const size_t csr = align_up((size_t)(nnz_cap + 8) * 4, 16);
  const size_t th = (size_t)K * F * F * 4;
  // FFMA2 variant: Theta travels with the double-buffered stage.  MMA variant: ONE Theta buffer with its own
  // barrier, refilled right after the last order's MMAs (the copy hides behind the epilogue and the next
  // graph's first propagation) -- 4 KB less per warp, 16 warps per SM.
  const size_t stage = t0 + (mma ? 0 : th) + 2 * csr;
  return align_up(32 + ta + (mma ? th : 0) + 2 * stage, 128);
}

static WarpCfg warp_config(int F, int K, int max_nodes) {
  WarpCfg c{0, 0, 0, 0, false, 0, 0, false};
  c.mma = (F == 16) && getenv("FETA_CHEB_NO_MMA") == nullptr;   // F = 8: one k-step, the FFMA2 path is faster
  if (!(F == 4 || F == 8 || F == 16) || max_nodes < 1 || max_nodes > 64) return c;
  if ((size_t)K * F * F * 4 > 16 * 1024) return c;
  c.rpl = max_nodes <= 32 ? 1 : 2;
  // buffers sized for the largest graph; an MMA tile may read up to 8 rows past it -- that lands in the
  // next region of the same warp (stage T0 / Theta), is never written and only feeds discarded output rows
  c.rows_cap = (max_nodes + 7) / 8 * 8;
  c.nnz_cap = c.rows_cap * 4;
  c.per_warp = warp_region_bytes(F, K, c.rows_cap, c.nnz_cap, c.mma);
  int w = (int)((227 * 1024) / c.per_warp);
  const int wmax = F >= 16 ? 12 : 16;
  if (w > wmax) w = wmax;
  if (w < 4) return c;
  c.warps = w;
  c.smem = c.per_warp * w;
  c.ok = true;
  return c;
}

template <int F>
__device__ __forceinline__ void apply_theta_s(float2 (&acc)[F / 2], const float (&t)[F], const float* __restrict__ th) {
#pragma unroll
  for (int i = 0; i < F; ++i) {
    const float2 tt = make_float2(t[i], t[i]);
#pragma unroll
    for (int q = 0; q < F / 4; ++q) {
      const float4 w = *reinterpret_cast<const float4*>(th + i * F + 4 * q);
      acc[2 * q] = __ffma2_rn(tt, make_float2(w.x, w.y), acc[2 * q]);
      acc[2 * q + 1] = __ffma2_rn(tt, make_float2(w.z, w.w), acc[2 * q + 1]);
    }
  }
}

template <int F, int LD, bool STAGED>
__device__ __forceinline__ void gather_row_w(float (&t)[F], const float* __restrict__ buf, int r0,
                                             const int32_t* __restrict__ ci, const float* __restrict__ cv, int e0,
                                             int e1) {
  float2 a2[F / 2];
#pragma unroll
  for (int i = 0; i < F / 2; ++i) a2[i] = make_float2(0.f, 0.f);
  for (int e = e0; e < e1; ++e) {
    const int c = (STAGED ? ci[e] : __ldg(ci + e)) - r0;
    const float w = STAGED ? cv[e] : __ldg(cv + e);
    const float2 ww = make_float2(w, w);
    const float* row = buf + c * LD;
#pragma unroll
    for (int q = 0; q < F / 4; ++q) {
      const float4 a = *reinterpret_cast<const float4*>(row + 4 * q);
      a2[2 * q] = __ffma2_rn(ww, make_float2(a.x, a.y), a2[2 * q]);
      a2[2 * q + 1] = __ffma2_rn(ww, make_float2(a.z, a.w), a2[2 * q + 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < F / 2; ++i) t[2 * i] = a2[i].x, t[2 * i + 1] = a2[i].y;
}

// ---- XOR-swizzled [rows, 16] fp32 slabs (MMA variant) ------------------------------------------------
// A row is 64 B = 4 chunks of 16 B; chunk q of row r is stored at chunk position q ^ ((r >> 1) & 3).  Eight
// consecutive rows then cover all 32 banks for a lane-per-row LDS.128/STS.128, and an m16n8k8 A-fragment
// load (8 rows x 4 consecutive words) is conflict-free as well; a dense slab has 4-way conflicts for both.
__device__ __forceinline__ int swz16(int r) { return ((r >> 1) & 3) << 2; }

template <bool STAGED>
__device__ __forceinline__ void gather_row_swz(float (&t)[16], const float* __restrict__ buf, int r0,
                                               const int32_t* __restrict__ ci, const float* __restrict__ cv, int e0,
                                               int e1) {
  float2 a2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a2[i] = make_float2(0.f, 0.f);
  if constexpr (STAGED) {
    // two edges per trip: 8 independent LDS.128 in flight, half the loop overhead.  The slot one past the
    // row's last edge is inside the staged slice (8 words of slack); its contents are replaced, not scaled.
    for (int e = e0; e < e1; e += 2) {
      const bool two = e + 1 < e1;
      const int cA = ci[e] - r0;
      const float wA = cv[e];
      int cB = ci[e + 1] - r0;
      float wB = cv[e + 1];
      cB = two ? cB : cA;
      wB = two ? wB : 0.0f;
      const float* rowA = buf + cA * 16;
      const float* rowB = buf + cB * 16;
      const int swA = swz16(cA), swB = swz16(cB);
      const float2 wwA = make_float2(wA, wA), wwB = make_float2(wB, wB);
      float4 a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a[q] = *reinterpret_cast<const float4*>(rowA + ((4 * q) ^ swA));
        b[q] = *reinterpret_cast<const float4*>(rowB + ((4 * q) ^ swB));
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        a2[2 * q] = __ffma2_rn(wwA, make_float2(a[q].x, a[q].y), a2[2 * q]);
        a2[2 * q + 1] = __ffma2_rn(wwA, make_float2(a[q].z, a[q].w), a2[2 * q + 1]);
        a2[2 * q] = __ffma2_rn(wwB, make_float2(b[q].x, b[q].y), a2[2 * q]);
        a2[2 * q + 1] = __ffma2_rn(wwB, make_float2(b[q].z, b[q].w), a2[2 * q + 1]);
      }
    }
  } else {
    for (int e = e0; e < e1; ++e) {
      const int c = __ldg(ci + e) - r0;
      const float w = __ldg(cv + e);
      const float2 ww = make_float2(w, w);
      const float* row = buf + c * 16;
      const int sw = swz16(c);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 a = *reinterpret_cast<const float4*>(row + ((4 * q) ^ sw));
        a2[2 * q] = __ffma2_rn(ww, make_float2(a.x, a.y), a2[2 * q]);
        a2[2 * q + 1] = __ffma2_rn(ww, make_float2(a.z, a.w), a2[2 * q + 1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) t[2 * i] = a2[i].x, t[2 * i + 1] = a2[i].y;
}

// ---- 3xTF32 on the (legacy, warp-level) tensor-core path: T_k . Theta_k as m16n8k8 MMAs -------------
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  // the tensor core reads only the top 19 bits of a tf32 operand (truncation), so "hi" is x itself and the
  // remainder is exact in fp32:  x = trunc19(x) + lo,  |lo| <= 2^-10 |x|;  lo is truncated again by the
  // hardware (relative error of the 3-term product ~ 2^-20).  2 instructions per element (LOP3 + FADD).
  hi = __float_as_uint(x);
  lo = __float_as_uint(x - __uint_as_float(hi & 0xFFFFE000u));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// acc[mt][nt] += T[16mt .. 16mt+15, :] . Theta_k   for mt < MT, with hi/lo splits of both operands
// Tb is an XOR-swizzled [rows, 16] slab (see swz16)
template <int F, int MTMAX>
__device__ __forceinline__ void mma_order(float (&acc)[MTMAX][F / 8][4], const float* __restrict__ Tb,
                                          const float* __restrict__ thk, int MT, int lane) {
  constexpr int NT = F / 8, KS = F / 8;
  const int g = lane >> 2, tq = lane & 3;
  uint32_t bh[NT][KS][2], bl[NT][KS][2];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      split_tf32(thk[(8 * ks + tq) * F + 8 * nt + g], bh[nt][ks][0], bl[nt][ks][0]);
      split_tf32(thk[(8 * ks + tq + 4) * F + 8 * nt + g], bh[nt][ks][1], bl[nt][ks][1]);
    }
#pragma unroll
  for (int mt = 0; mt < MTMAX; ++mt) {
    if (mt < MT) {
      static_assert(F == 16, "swizzled slabs are 16 floats wide");
      const float* r0 = Tb + (16 * mt + g) * F + tq;   // rows 16mt+g and +8 share the swizzle ((g >> 1) & 3)
      const float* r1 = r0 + 8 * F;
      const int sw = swz16(g);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ah[4], al[4];
        const int c0 = (8 * ks) ^ sw, c1 = (8 * ks + 4) ^ sw;
        split_tf32(r0[c0], ah[0], al[0]);
        split_tf32(r1[c0], ah[1], al[1]);
        split_tf32(r0[c1], ah[2], al[2]);
        split_tf32(r1[c1], ah[3], al[3]);
        // the three split terms are issued nt-interleaved: consecutive MMAs hit different accumulators
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

template <int RPL>
struct GraphDesc {  // scalars of one graph, fetched ahead of use
  int r0, r1, e_lo, e_hi;
  int e0[RPL], e1[RPL];
};

template <int F, bool USE_MMA>
constexpr int warp_kernel_max_threads() {
  // FFMA2 variant at F = 16: at most 12 warps fit shared memory -> ~170 registers/thread allowed
  return F >= 16 ? 384 : 512;   // 12 warps = 3 per scheduler -> 168 registers; 13..16 warps would cap at 128 (spills, measured slower)
}

template <int F, int RPL, bool USE_MMA>
__global__ void __launch_bounds__(warp_kernel_max_threads<F, USE_MMA>(), 1) cheb_fwd_warp_kernel(
    const float* __restrict__ x, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
    const float* __restrict__ vals, const int32_t* __restrict__ graph_ptr, const float* __restrict__ theta,
    int64_t sk, int64_t sg, const float* __restrict__ bias, float* __restrict__ out, int64_t R, int64_t G, int K,
    int nnz_cap, int rows_cap, int per_warp_bytes, int32_t* meta, int max_nodes) {
  constexpr int LD = F + 4;
  const uint32_t TA = (uint32_t)rows_cap * (USE_MMA ? F : LD) * 4;  // odd-order buffer (padded / swizzled)
  const uint32_t TB = (uint32_t)rows_cap * F * 4;   // dense x slab / even-order buffer (per stage)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // device-side plan guard (same contract as the chunk kernel)
  if (meta != nullptr) {
    const bool ok = meta[FETA_META_SORTED] == 1 && meta[FETA_META_BLOCKDIAG] == 1 &&
                    meta[FETA_META_NUM_GRAPHS] == (int32_t)G && meta[FETA_META_MAX_NODES] <= max_nodes &&
                    meta[FETA_META_BAD_INDEX] == 0;
    if (!ok) {
      if (threadIdx.x == 0) meta[FETA_META_GUARD] = 1;
      nan_fill(out, R * F);     // a refused launch leaves NaN, never uninitialised memory
      return;
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int64_t gw = (int64_t)blockIdx.x * warps + warp, stride = (int64_t)gridDim.x * warps;
  if (gw >= G) return;
  const int n_it = (int)((G - gw + stride - 1) / stride);

  unsigned char* base = smem_raw + (size_t)warp * per_warp_bytes;
  const uint32_t csr_bytes = (uint32_t)(((nnz_cap + 8) * 4 + 15) / 16 * 16);
  const uint32_t th_bytes = (uint32_t)K * F * F * 4;
  const uint32_t th_stage = USE_MMA ? 0u : th_bytes;   // Theta bytes inside a stage (FFMA2 variant only)
  const uint32_t stage_bytes = TB + th_stage + 2 * csr_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(base);   // [0], [1]: stages; [2]: the single Theta buffer
  float* bufA = reinterpret_cast<float*>(base + 32);
  unsigned char* th_single = base + 32 + TA;            // MMA variant
  unsigned char* stage0 = base + 32 + TA + (USE_MMA ? th_bytes : 0u);
  const bool theta_contig = (sk == (int64_t)F * F);

  if (lane == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int nnz_total = __ldg(rowptr + R);
  float bias_r[F];
#pragma unroll
  for (int i = 0; i < F; ++i) bias_r[i] = bias ? __ldg(bias + i) : 0.0f;
  float bias_m[(F + 7) / 8][2];  // MMA variant: the two output columns this lane owns in every n-tile
#pragma unroll
  for (int nt = 0; nt < (USE_MMA ? F / 8 : 0); ++nt) {
    bias_m[nt][0] = bias ? __ldg(bias + 8 * nt + 2 * (lane & 3)) : 0.0f;
    bias_m[nt][1] = bias ? __ldg(bias + 8 * nt + 2 * (lane & 3) + 1) : 0.0f;
  }

  auto fetch_a = [&](int it, GraphDesc<RPL>& d) {  // stage "a": row range of the graph
    d.r0 = d.r1 = 0;
    if (it < n_it) {
      const int64_t g = gw + (int64_t)it * stride;
      d.r0 = __ldg(graph_ptr + g);
      d.r1 = __ldg(graph_ptr + g + 1);
    }
  };
  auto fetch_b = [&](int it, GraphDesc<RPL>& d) {  // stage "b": CSR extents (depends on stage a)
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
    return (a_hi - a_lo <= nnz_cap + 4) && (a_hi <= nnz_total);
  };
  auto issue = [&](int it, const GraphDesc<RPL>& d) {  // stage "c": TMA copies into stage[it & 1]
    if (it >= n_it) return;
    const int s = it & 1;
    unsigned char* st = stage0 + (size_t)s * stage_bytes;
    const uint32_t bar = smem_u32(&bars[s]);
    const int n = d.r1 - d.r0;
    const int64_t g = gw + (int64_t)it * stride;
    int a_lo, a_hi;
    const bool staged = staged_csr(d, a_lo, a_hi);
    const bool copy_csr = staged && (d.e_hi > d.e_lo);
    const uint32_t bytes = (uint32_t)n * F * 4 + th_stage + (copy_csr ? 2u * (uint32_t)(a_hi - a_lo) * 4u : 0u);
    fence_proxy_async();  // generic-proxy writes into this stage (even orders) before the TMA refills it
    __syncwarp();         // every lane is done with this stage (used two graphs ago)
    if (lane == 0) {  // one thread drives the TMA: 1 x-slab + 1 (or K) Theta + 2 CSR copies per graph
      mbar_arrive_expect_tx(bar, bytes);
      bulk_g2s(smem_u32(st), x + (size_t)d.r0 * F, (uint32_t)n * F * 4, bar);
      if constexpr (!USE_MMA) {
        if (theta_contig) {
          bulk_g2s(smem_u32(st + TB), theta + g * sg, th_bytes, bar);
        } else {
          for (int k = 0; k < K; ++k)
            bulk_g2s(smem_u32(st + TB + (size_t)k * F * F * 4), theta + g * sg + (int64_t)k * sk, F * F * 4, bar);
        }
      }
      if (copy_csr) {
        bulk_g2s(smem_u32(st + TB + th_stage), colidx + a_lo, (uint32_t)(a_hi - a_lo) * 4, bar);
        bulk_g2s(smem_u32(st + TB + th_stage + csr_bytes), vals + a_lo, (uint32_t)(a_hi - a_lo) * 4, bar);
      }
    }
  };
  // MMA variant: refill the single Theta buffer for graph `it`; the caller has synchronised the warp after
  // the last read of the previous graph's Theta
  auto issue_theta = [&](int it) {
    if (it >= n_it) return;
    if (lane == 0) {
      const int64_t g = gw + (int64_t)it * stride;
      const uint32_t bar = smem_u32(&bars[2]);
      mbar_arrive_expect_tx(bar, th_bytes);
      if (theta_contig) {
        bulk_g2s(smem_u32(th_single), theta + g * sg, th_bytes, bar);
      } else {
        for (int k = 0; k < K; ++k)
          bulk_g2s(smem_u32(th_single + (size_t)k * F * F * 4), theta + g * sg + (int64_t)k * sk, F * F * 4, bar);
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
  if constexpr (USE_MMA) issue_theta(0);

  for (int it = 0; it < n_it; ++it) {
    fetch_a(it + 3, d3);
    fetch_b(it + 2, d2);
    issue(it + 1, d1);

    // ---------------- compute graph `it` out of stage[it & 1]
    const int s = it & 1;
    unsigned char* st = stage0 + (size_t)s * stage_bytes;
    float* T0 = reinterpret_cast<float*>(st);
    const float* th = reinterpret_cast<const float*>(USE_MMA ? th_single : st + TB);
    int a_lo, a_hi;
    const bool staged = staged_csr(d0, a_lo, a_hi);
    // staged CSR slices stay typed as shared-memory pointers (a select against the global arrays would turn
    // every access into a generic load)
    const int32_t* ci_s = reinterpret_cast<const int32_t*>(st + TB + th_stage) - a_lo;
    const float* cv_s = reinterpret_cast<const float*>(st + TB + th_stage + csr_bytes) - a_lo;
    const int n = d0.r1 - d0.r0;
    mbar_wait(smem_u32(&bars[s]), (uint32_t)((it >> 1) & 1));
    if (lane == 0) bulk_wait_read0();  // the previous graph's output store has drained bufA
    __syncwarp();

    if constexpr (USE_MMA) {
      // ---- tensor-core variant: the filter application T_k . Theta_k runs as 3xTF32 m16n8k8 MMAs
      constexpr int MTMAX = 2 * RPL, NT = F / 8;
      float acc[MTMAX][NT][4];
#pragma unroll
      for (int mt = 0; mt < MTMAX; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[mt][nt][q] = 0.0f;
      const int MT = (n + 15) >> 4;
      {  // re-lay the bulk-copied (dense) x slab into the swizzled layout, in place: the permutation stays
         // inside a row and a pass of 32 chunks covers 8 whole rows
        float4 v[4 * RPL];
        float4* T4 = reinterpret_cast<float4*>(T0);
#pragma unroll
        for (int p = 0; p < 4 * RPL; ++p) {
          const int i = 32 * p + lane;
          if (i < 4 * n) v[p] = T4[i];
        }
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 4 * RPL; ++p) {
          const int i = 32 * p + lane;
          if (i < 4 * n) T4[i ^ ((i >> 3) & 3)] = v[p];
        }
        __syncwarp();
      }
      // order of work: T_1 is propagated BEFORE the first filter application, so that the Theta copy (issued
      // at the end of the previous graph) has the epilogue, the re-lay and one propagation to land
      float t[F];
      for (int k = 1; k < K; ++k) {
        // even orders live in the stage buffer T0, odd orders in bufA (both swizzled); T_k overwrites the
        // own row of T_{k-2} (nobody else reads it any more)
        const bool odd = (k & 1) != 0;
        const float* src = odd ? T0 : bufA;
        float* dst = odd ? bufA : T0;
#pragma unroll
        for (int m = 0; m < RPL; ++m) {
          const int row = lane + 32 * m;
          if (row < n) {
            if (staged) gather_row_swz<true>(t, src, d0.r0, ci_s, cv_s, d0.e0[m], d0.e1[m]);
            else gather_row_swz<false>(t, src, d0.r0, colidx, vals, d0.e0[m], d0.e1[m]);
            float* drow = dst + row * F;
            const int sw = swz16(row);
            if (k >= 2) {
#pragma unroll
              for (int q = 0; q < F / 4; ++q) {
                const float4 o = *reinterpret_cast<const float4*>(drow + ((4 * q) ^ sw));
                t[4 * q] = fmaf(2.0f, t[4 * q], -o.x), t[4 * q + 1] = fmaf(2.0f, t[4 * q + 1], -o.y);
                t[4 * q + 2] = fmaf(2.0f, t[4 * q + 2], -o.z), t[4 * q + 3] = fmaf(2.0f, t[4 * q + 3], -o.w);
              }
            }
#pragma unroll
            for (int q = 0; q < F / 4; ++q)
              *reinterpret_cast<float4*>(drow + ((4 * q) ^ sw)) =
                  make_float4(t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]);
          }
        }
        __syncwarp();
        if (k == 1) {
          mbar_wait(smem_u32(&bars[2]), (uint32_t)(it & 1));
          mma_order<F, MTMAX>(acc, T0, th, MT, lane);   // k = 0 (before order 2 overwrites the x slab)
        }
        mma_order<F, MTMAX>(acc, dst, th + (size_t)k * F * F, MT, lane);
      }
      if (K == 1) {
        mbar_wait(smem_u32(&bars[2]), (uint32_t)(it & 1));
        mma_order<F, MTMAX>(acc, T0, th, MT, lane);
      }
      __syncwarp();          // every lane has read its last Theta fragment
      issue_theta(it + 1);
      __syncwarp();   // all lanes are done reading bufA before it becomes the output staging slab
      const int g = lane >> 2, tq = lane & 3;
#pragma unroll
      for (int mt = 0; mt < MTMAX; ++mt) {
        if (mt < MT) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const int col = 8 * nt + 2 * tq;
            const float b0 = bias_m[nt][0], b1 = bias_m[nt][1];
            const int ra = 16 * mt + g, rb = ra + 8;
            if (ra < n) *reinterpret_cast<float2*>(bufA + ra * F + col) = make_float2(acc[mt][nt][0] + b0, acc[mt][nt][1] + b1);
            if (rb < n) *reinterpret_cast<float2*>(bufA + rb * F + col) = make_float2(acc[mt][nt][2] + b0, acc[mt][nt][3] + b1);
          }
        }
      }
    } else {
    float2 acc[RPL][F / 2];
      float t[F];
  #pragma unroll
      for (int m = 0; m < RPL; ++m) {
  #pragma unroll
        for (int i = 0; i < F / 2; ++i) acc[m][i] = make_float2(0.f, 0.f);
        const int row = lane + 32 * m;
        if (row < n) {
  #pragma unroll
          for (int q = 0; q < F / 4; ++q) {
            const float4 a = *reinterpret_cast<const float4*>(T0 + row * F + 4 * q);
            t[4 * q] = a.x, t[4 * q + 1] = a.y, t[4 * q + 2] = a.z, t[4 * q + 3] = a.w;
          }
          apply_theta_s<F>(acc[m], t, th);  // k = 0
        }
      }
      // even orders live in the (dense) stage buffer, odd orders in the padded bufA; T_k overwrites the
      // own row of T_{k-2} (nobody else reads it any more)
      for (int k = 1; k < K; ++k) {
        const bool odd = (k & 1) != 0;
  #pragma unroll
        for (int m = 0; m < RPL; ++m) {
          const int row = lane + 32 * m;
          if (row < n) {
            if (odd) {
              if (staged) gather_row_w<F, F, true>(t, T0, d0.r0, ci_s, cv_s, d0.e0[m], d0.e1[m]);
              else gather_row_w<F, F, false>(t, T0, d0.r0, colidx, vals, d0.e0[m], d0.e1[m]);
            } else {
              if (staged) gather_row_w<F, LD, true>(t, bufA, d0.r0, ci_s, cv_s, d0.e0[m], d0.e1[m]);
              else gather_row_w<F, LD, false>(t, bufA, d0.r0, colidx, vals, d0.e0[m], d0.e1[m]);
            }
            float* drow = odd ? bufA + row * LD : T0 + row * F;
            if (k >= 2) {
  #pragma unroll
              for (int q = 0; q < F / 4; ++q) {
                const float4 o = *reinterpret_cast<const float4*>(drow + 4 * q);
                t[4 * q] = fmaf(2.0f, t[4 * q], -o.x), t[4 * q + 1] = fmaf(2.0f, t[4 * q + 1], -o.y);
                t[4 * q + 2] = fmaf(2.0f, t[4 * q + 2], -o.z), t[4 * q + 3] = fmaf(2.0f, t[4 * q + 3], -o.w);
              }
            }
            if (k + 1 < K) {
  #pragma unroll
              for (int q = 0; q < F / 4; ++q)
                *reinterpret_cast<float4*>(drow + 4 * q) = make_float4(t[4 * q], t[4 * q + 1], t[4 * q + 2], t[4 * q + 3]);
            }
            apply_theta_s<F>(acc[m], t, th + (size_t)k * F * F);
          }
        }
        __syncwarp();
      }
      // epilogue: + bias, dense slab in bufA, one bulk store
  #pragma unroll
      for (int m = 0; m < RPL; ++m) {
        const int row = lane + 32 * m;
        if (row < n) {
  #pragma unroll
          for (int q = 0; q < F / 4; ++q)
            *reinterpret_cast<float4*>(bufA + row * F + 4 * q) =
                make_float4(acc[m][2 * q].x + bias_r[4 * q], acc[m][2 * q].y + bias_r[4 * q + 1],
                            acc[m][2 * q + 1].x + bias_r[4 * q + 2], acc[m][2 * q + 1].y + bias_r[4 * q + 3]);
        }
      }

    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      bulk_s2g(out + (size_t)d0.r0 * F, smem_u32(bufA), (uint32_t)n * F * 4);
      bulk_commit();
    }
    d0 = d1;
    d1 = d2;
    d2 = d3;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
}

template <int F, int RPL, bool USE_MMA>
static int launch_warp(const WarpCfg& c, const float* x, const int32_t* rowptr, const int32_t* colidx,
                       const float* vals, const int32_t* graph_ptr, const float* theta, int64_t sk, int64_t sg,
                       const float* bias, float* out, int64_t R, int64_t G, int K, int32_t* meta, int max_nodes,
                       cudaStream_t st) {
  FETA_CUDA(cudaFuncSetAttribute(cheb_fwd_warp_kernel<F, RPL, USE_MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)c.smem));
  int64_t grid = ceil_div(G, c.warps);
  if (grid > kNumSMs) grid = kNumSMs;
  cheb_fwd_warp_kernel<F, RPL, USE_MMA><<<(unsigned)grid, c.warps * 32, c.smem, st>>>(
      x, rowptr, colidx, vals, graph_ptr, theta, sk, sg, bias, out, R, G, K, c.nnz_cap, c.rows_cap, (int)c.per_warp,
      meta, max_nodes);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

// returns FETA_OK if launched, 1 if this shape is not eligible (caller falls back to the chunk kernel)
int cheb_fwd_warp_try(const float* x, const int32_t* rowptr, const int32_t* colidx, const float* vals,
                      const int32_t* graph_ptr, const float* theta, int64_t sk, int64_t sg, const float* bias,
                      float* out, int64_t R, int64_t G, int K, int F, int max_nodes, int32_t* meta, cudaStream_t st) {
  WarpCfg c = warp_config(F, K, max_nodes);
  if (!c.ok) return 1;
  if (((uintptr_t)colidx % 16) || ((uintptr_t)vals % 16)) return 1;
#define FETA_WARP_CASE(F_, R_, M_)                                                                                  \
  if (F == F_ && c.rpl == R_ && c.mma == M_)                                                                        \
    return launch_warp<F_, R_, M_>(c, x, rowptr, colidx, vals, graph_ptr, theta, sk, sg, bias, out, R, G, K, meta, \
                                   max_nodes, st);
  FETA_WARP_CASE(4, 1, false) FETA_WARP_CASE(4, 2, false) FETA_WARP_CASE(8, 1, false) FETA_WARP_CASE(8, 2, false)
  FETA_WARP_CASE(16, 1, false) FETA_WARP_CASE(16, 2, false)
  FETA_WARP_CASE(16, 1, true) FETA_WARP_CASE(16, 2, true)
#undef FETA_WARP_CASE
  return 1;
}

}  // namespace feta
