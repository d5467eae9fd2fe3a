// A6: kernel-biased attention forward on the 5th-gen tensor cores (tcgen05 + TMEM), fp32-grade
// accuracy through a 3xTF32 split.  Same contract as attn_fwd_kernel in attention.cu.
//
// One CTA (128 threads) per (graph b, head h, tile of 128 query rows):
//   1. Q (pre-scaled), K and V^T of the (b, h) are staged in shared memory in the UMMA canonical
//      K-major / no-swizzle layout ([row/8][k/4][row%8][k%4], 8x16-byte core matrices), each as a
//      TF32 "hi" part and an fp32 remainder "lo" part.
//   2. S = Q K^T:  tcgen05.mma.kind::tf32 (M=128, N=NK, K=8) x 3 terms (hi.hi + hi.lo + lo.hi),
//      accumulators in TMEM (one lane per query row, one column per key).
//   3. Softmax-like normalisation with the kernel bias, one thread per query row straight out of
//      TMEM (tcgen05.ld 32x32b): key-padding mask, row max, exp, * pe, row sum, divide, write the
//      attention row; P goes back INTO TMEM (tcgen05.st) as hi / lo parts.
//   4. O = P V:  tcgen05.mma with the A operand read from TMEM (P) and B = V^T from shared memory,
//      again 3 terms; O is read back with tcgen05.ld and written per head.
// Only the two contractions run on the tensor cores; everything else is fp32 CUDA-core code.
#include <math.h>

#include "common.cuh"
#include "umma.cuh"

namespace feta {



constexpr int kTcThreads = 128;

template <int DH>
__global__ void __launch_bounds__(kTcThreads) attn_fwd_tc_kernel(
    const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v, int64_t sn, int64_t sb,
    const float* __restrict__ pe, const uint8_t* __restrict__ mask, float* __restrict__ attn,
    float* __restrict__ o_heads, int64_t osn, int64_t osb, float* __restrict__ rowflag, int H, int nmax,
    float scale) {
  using namespace tc;
  constexpr int NO = DH < 16 ? 16 : DH;     // N of the P.V product (multiple of 16 for M = 128)
  constexpr int KC = DH / 4;                // 16-byte K-cores along the head dimension
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ uint32_t s_tmem;
  __shared__ int s_neff;

  const int bh = blockIdx.y, b = bh / H, h = bh - b * H;
  const int i0 = blockIdx.x * 128;
  const uint8_t* mk = mask + (size_t)b * nmax;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // n_eff (last un-masked key + 1) -> NK = keys rounded up to 32
  if (tid == 0) s_neff = 0;
  __syncthreads();
  {
    int loc = 0;
    for (int j = tid; j < nmax; j += kTcThreads)
      if (mk[j] == 0) loc = j + 1;
    if (loc > 0) atomicMax(&s_neff, loc);
  }
  __syncthreads();
  const int n = s_neff;
  const int NK = n > 0 ? ((n + 31) / 32) * 32 : 32;
  const int kcV = NK / 4;

  // shared-memory carve: Q hi/lo [128 x DH], K hi/lo [NK x DH], V^T hi/lo [NO x NK]
  const uint32_t q_bytes = 128 * DH * 4, k_bytes = (uint32_t)NK * DH * 4, v_bytes = (uint32_t)NO * NK * 4;
  unsigned char* sQh = smem_raw;
  unsigned char* sQl = sQh + q_bytes;
  unsigned char* sKh = sQl + q_bytes;
  unsigned char* sKl = sKh + k_bytes;
  unsigned char* sVh = sKl + k_bytes;
  unsigned char* sVl = sVh + v_bytes;

  // TMEM: [0, NK) S then P_hi, [NK, 2NK) P_lo, [2NK, 2NK + NO) O
  uint32_t ncols = 32;
  while ((int)ncols < 2 * NK + NO) ncols <<= 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // ---- stage operands (hi / lo split), zero padding rows and keys
  const int64_t base = (int64_t)b * sb + h * DH;
  for (int idx = tid; idx < 128 * DH; idx += kTcThreads) {
    const int r = idx / DH, c = idx - r * DH;
    const int i = i0 + r;
    float x = 0.0f;
    if (i < nmax && mk[i] == 0) x = __ldg(q + base + (int64_t)i * sn + c) * scale;   // q * scaling first
    const float hi = tf32_hi(x);
    const uint32_t off = canon(r, c, KC);
    *reinterpret_cast<float*>(sQh + off) = hi;
    *reinterpret_cast<float*>(sQl + off) = x - hi;
  }
  for (int idx = tid; idx < NK * DH; idx += kTcThreads) {
    const int j = idx / DH, c = idx - j * DH;
    float kk = 0.0f, vv = 0.0f;
    if (j < n && mk[j] == 0) {
      kk = __ldg(k + base + (int64_t)j * sn + c);
      vv = __ldg(v + base + (int64_t)j * sn + c);
    }
    const float kh = tf32_hi(kk), vh = tf32_hi(vv);
    const uint32_t offk = canon(j, c, KC);
    *reinterpret_cast<float*>(sKh + offk) = kh;
    *reinterpret_cast<float*>(sKl + offk) = kk - kh;
    const uint32_t offv = canon(c, j, kcV);            // V^T: row = channel, K = key
    *reinterpret_cast<float*>(sVh + offv) = vh;
    *reinterpret_cast<float*>(sVl + offv) = vv - vh;
  }
  if (NO > DH) {                                         // zero the padding channels of V^T
    for (int idx = tid; idx < (NO - DH) * NK; idx += kTcThreads) {
      const int c = DH + idx / NK, j = idx % NK;
      const uint32_t offv = canon(c, j, kcV);
      *reinterpret_cast<float*>(sVh + offv) = 0.0f;
      *reinterpret_cast<float*>(sVl + offv) = 0.0f;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t tS = tmem, tPl = tmem + (uint32_t)NK, tO = tmem + 2u * (uint32_t)NK;

  // ---- S = Q K^T  (3xTF32)
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, NK);
    const uint32_t sbo = KC * 128;
#pragma unroll
    for (int ks = 0; ks < DH / 8; ++ks) {
      const uint64_t aH = make_desc(smem_u32(sQh) + ks * 256, 128, sbo), aL = make_desc(smem_u32(sQl) + ks * 256, 128, sbo);
      const uint64_t bH = make_desc(smem_u32(sKh) + ks * 256, 128, sbo), bL = make_desc(smem_u32(sKl) + ks * 256, 128, sbo);
      mma_ss(tS, aH, bH, idesc, ks > 0);
      mma_ss(tS, aH, bL, idesc, 1);
      mma_ss(tS, aL, bH, idesc, 1);
    }
    mma_commit(smem_u32(&s_bar[0]));
  }
  mbar_wait_parity(smem_u32(&s_bar[0]), 0);
  fence_after();

  // ---- normalisation, one thread per query row
  const int r = tid, i = i0 + r;
  const bool live = (i < nmax) && (mk[i] == 0);
  const uint32_t lane_base = ((uint32_t)(warp * 32)) << 16;
  const float* perow = (pe != nullptr && i < nmax) ? pe + ((size_t)b * nmax + i) * nmax : nullptr;
  float* arow = (i < nmax) ? attn + (((size_t)b * H + h) * nmax + i) * nmax : nullptr;
  float vals[16];
  float m = -INFINITY;
  for (int c0 = 0; c0 < NK; c0 += 16) {
    tmem_ld16(tS + lane_base + c0, vals);
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int j = c0 + t;
      if (j < n && mk[j] == 0) m = fmaxf(m, vals[t]);
    }
  }
  float sum = 0.0f;
  for (int c0 = 0; c0 < NK; c0 += 16) {
    tmem_ld16(tS + lane_base + c0, vals);
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int j = c0 + t;
      float e = 0.0f;
      if (live && j < n && mk[j] == 0) {
        e = expf(vals[t] - m);
        if (perow) e *= __ldg(perow + j);
      }
      vals[t] = e;
      sum += e;
    }
    tmem_st16(tS + lane_base + c0, vals);
  }
  wait_st();
  const float denom = fmaxf(sum, 1e-6f);
  for (int c0 = 0; c0 < NK; c0 += 16) {
    tmem_ld16(tS + lane_base + c0, vals);
    float lo[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const int j = c0 + t;
      const float p = vals[t] / denom;
      if (arow != nullptr && j < nmax) arow[j] = p;
      const float hi = tf32_hi(p);
      vals[t] = hi;
      lo[t] = p - hi;
    }
    tmem_st16(tS + lane_base + c0, vals);
    tmem_st16(tPl + lane_base + c0, lo);
  }
  if (arow != nullptr)
    for (int j = NK; j < nmax; ++j) arow[j] = 0.0f;       // keys beyond the padded tile
  if (i < nmax) rowflag[((size_t)b * H + h) * nmax + i] = (live && sum > 1e-6f) ? 1.0f : 0.0f;
  wait_st();
  fence_before();
  __syncthreads();
  fence_after();

  // ---- O = P V  (A = P from TMEM, B = V^T from shared memory; 3xTF32)
  if (tid == 0) {
    const uint32_t idesc = make_idesc(128, NO);
    const uint32_t sbo = (uint32_t)kcV * 128;
    for (int kk = 0; kk < NK / 8; ++kk) {
      const uint64_t bH = make_desc(smem_u32(sVh) + kk * 256, 128, sbo), bL = make_desc(smem_u32(sVl) + kk * 256, 128, sbo);
      mma_ts(tO, tS + kk * 8, bH, idesc, kk > 0);
      mma_ts(tO, tS + kk * 8, bL, idesc, 1);
      mma_ts(tO, tPl + kk * 8, bH, idesc, 1);
    }
    mma_commit(smem_u32(&s_bar[1]));
  }
  mbar_wait_parity(smem_u32(&s_bar[1]), 0);
  fence_after();
  tmem_ld16(tO + lane_base, vals);
  if (i < nmax) {
    float* orow = o_heads + (int64_t)i * osn + (int64_t)b * osb + h * DH;
#pragma unroll
    for (int c = 0; c < (DH < 16 ? DH : 16); ++c) orow[c] = live ? vals[c] : 0.0f;
  }
  if (DH > 16) {
    tmem_ld16(tO + lane_base + 16, vals);
    if (i < nmax) {
      float* orow = o_heads + (int64_t)i * osn + (int64_t)b * osb + h * DH;
#pragma unroll
      for (int c = 0; c < 16; ++c)
        if (16 + c < DH) orow[16 + c] = live ? vals[c] : 0.0f;
    }
  }
  fence_before();
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols) : "memory");
}

static size_t attn_tc_smem(int dh, int nmax) {
  const int NK = ((nmax + 31) / 32) * 32;
  const int NO = dh < 16 ? 16 : dh;
  return (size_t)2 * 128 * dh * 4 + (size_t)2 * NK * dh * 4 + (size_t)2 * NO * NK * 4 + 128;
}

// returns FETA_OK if launched, 1 if the shape is not eligible (caller uses the CUDA-core kernel)
int attn_fwd_tc_try(const float* q, const float* k, const float* v, int64_t sn, int64_t sb, const float* pe,
                    const uint8_t* mask, float* attn, float* o_heads, int64_t osn, int64_t osb, float* rowflag, int B,
                    int H, int nmax, int dh, float scale, cudaStream_t st) {
  if (!(dh == 8 || dh == 16 || dh == 32)) return 1;
  const int NK = ((nmax + 31) / 32) * 32;
  const int NO = dh < 16 ? 16 : dh;
  if (2 * NK + NO > 512) return 1;
  const size_t smem = attn_tc_smem(dh, nmax);
  if (smem > 200 * 1024) return 1;
  dim3 grid((unsigned)ceil_div(nmax, 128), (unsigned)(B * H));
#define FETA_TC_CASE(D_)                                                                                              \
  if (dh == D_) {                                                                                                     \
    FETA_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel<D_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    attn_fwd_tc_kernel<D_><<<grid, kTcThreads, smem, st>>>(q, k, v, sn, sb, pe, mask, attn, o_heads, osn, osb, rowflag, \
                                                           H, nmax, scale);                                           \
    FETA_LAUNCH_CHECK();                                                                                              \
    return FETA_OK;                                                                                                   \
  }
  FETA_TC_CASE(8) FETA_TC_CASE(16) FETA_TC_CASE(32)
#undef FETA_TC_CASE
  return 1;
}

}  // namespace feta
