// A6 layer glue, BatchNorm variant -- the reference's ZINC default (experiments/run_transformer_gengcn.py:57,64:
// batch-norm unless --layer-norm): y = BatchNorm1d(a + bscale * b) over the flattened [Nmax * B, d] rows,
// INCLUDING the rows that are padding (the layer flattens the padded tensor before the norm).
//
// Training forward, 2 launches:
//   bn_stats_kernel   z = a + bscale*b written once; per-CTA Welford partials (count, mean, M2) per channel over the
//                     CTA's rows, rows weighted 0/1 by `roww` (static-shape batches are padded beyond the batch
//                     maximum the reference pads to: those rows must not enter the statistics);
//   bn_apply_kernel   every CTA merges the partials (Chan's formula, fixed order: deterministic), CTA 0 also writes
//                     mean / rstd for the backward pass and updates running_mean / running_var (unbiased) /
//                     num_batches_tracked; y = (z - mean) * rstd * gamma + beta.
// Backward, 2 launches: per-CTA partial sums of dy and dy*xhat, then dz = gamma*rstd*(dy - mean(dy) - xhat*mean(dy*xhat))
// (weighted rows only), dgamma / dbeta from the merged partials, d(b) = bscale * dz.
#include "common.cuh"

namespace feta {

constexpr int kBnThreads = 256;
constexpr int kBnRowsPerCta = 64;
constexpr int kBnMaxD = 256;

static inline int bn_blocks(int64_t T) {
  int64_t b = ceil_div(T > 0 ? T : 1, kBnRowsPerCta);
  return (int)(b < 4 * kNumSMs ? b : 4 * kNumSMs);
}

// thread = (channel c = tid % D, row slot rs = tid / D); requires D <= 256 and 256 % D == 0 or handled by stride loop
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                              const float* __restrict__ bscale,
                                                              const float* __restrict__ roww, float* __restrict__ z,
                                                              float* __restrict__ partial /* [nblk][3][D] */, int64_t T,
                                                              int D) {
  __shared__ float s_cnt[kBnThreads], s_mean[kBnThreads], s_m2[kBnThreads];
  const int slots = kBnThreads / D;                 // row slots per CTA pass
  const int c = threadIdx.x % D, rs = threadIdx.x / D;
  const int64_t rows_per = (T + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * rows_per, hi = lo + rows_per < T ? lo + rows_per : T;
  float cnt = 0.f, mean = 0.f, m2 = 0.f;
  if (rs < slots) {
    for (int64_t r = lo + rs; r < hi; r += slots) {
      float v = a[r * D + c];
      if (b) v += (bscale ? bscale[r] : 1.0f) * b[r * D + c];
      z[r * D + c] = v;
      const float w = roww ? roww[r] : 1.0f;
      if (w != 0.0f) {                              // Welford update
        cnt += 1.0f;
        const float d = v - mean;
        mean += d / cnt;
        m2 = fmaf(d, v - mean, m2);
      }
    }
  }
  s_cnt[threadIdx.x] = cnt, s_mean[threadIdx.x] = mean, s_m2[threadIdx.x] = m2;
  __syncthreads();
  if (threadIdx.x < D) {                            // merge the row slots of this channel in slot order
    float n = s_cnt[c], mu = s_mean[c], q = s_m2[c];
    for (int s = 1; s < slots; ++s) {
      const float nb = s_cnt[s * D + c], mb = s_mean[s * D + c], qb = s_m2[s * D + c];
      if (nb > 0.f) {
        const float tot = n + nb, d = mb - mu;
        mu += d * (nb / tot);
        q += qb + d * d * (n * nb / tot);
        n = tot;
      }
    }
    float* p = partial + (size_t)blockIdx.x * 3 * D;
    p[c] = n, p[D + c] = mu, p[2 * D + c] = q;
  }
}

__device__ __forceinline__ void bn_merge(const float* __restrict__ partial, int nblk, int D, int c, float& n, float& mu,
                                         float& q) {
  n = 0.f, mu = 0.f, q = 0.f;
  for (int k = 0; k < nblk; ++k) {
    const float* p = partial + (size_t)k * 3 * D;
    const float nb = p[c], mb = p[D + c], qb = p[2 * D + c];
    if (nb > 0.f) {
      const float tot = n + nb, d = mb - mu;
      mu += d * (nb / tot);
      q += qb + d * d * (n * nb / tot);
      n = tot;
    }
  }
}

__global__ void __launch_bounds__(kBnThreads) bn_apply_kernel(const float* __restrict__ z, const float* __restrict__ partial,
                                                              int nblk, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, float* __restrict__ y,
                                                              float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                              float* __restrict__ running_mean,
                                                              float* __restrict__ running_var,
                                                              int64_t* __restrict__ num_batches, float momentum, float eps,
                                                              int64_t T, int D) {
  __shared__ float s_scale[kBnMaxD], s_shift[kBnMaxD];
  if (threadIdx.x < D) {
    const int c = threadIdx.x;
    float n, mu, q;
    bn_merge(partial, nblk, D, c, n, mu, q);
    const float var = n > 0.f ? q / n : 0.f;
    const float rstd = rsqrtf(var + eps);
    const float g = gamma ? gamma[c] : 1.0f, bt = beta ? beta[c] : 0.0f;
    s_scale[c] = rstd * g;
    s_shift[c] = bt - mu * rstd * g;
    if (blockIdx.x == 0) {
      mean_out[c] = mu;
      rstd_out[c] = rstd;
      if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mu;
      if (running_var) running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (n > 1.f ? q / (n - 1.f) : var);
      if (num_batches && c == 0) num_batches[0] += 1;
    }
  }
  __syncthreads();
  const int64_t total = T * D;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < total; i += (int64_t)gridDim.x * blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(z + i);
    const int c = (int)(i % D);
    float4 o;
    o.x = fmaf(v.x, s_scale[c], s_shift[c]), o.y = fmaf(v.y, s_scale[c + 1], s_shift[c + 1]);
    o.z = fmaf(v.z, s_scale[c + 2], s_shift[c + 2]), o.w = fmaf(v.w, s_scale[c + 3], s_shift[c + 3]);
    *reinterpret_cast<float4*>(y + i) = o;
  }
}

// backward partials: per CTA, per channel: sum(dy), sum(dy * xhat), count   (weighted rows only)
__global__ void __launch_bounds__(kBnThreads) bn_bwd_stats_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ rstd,
                                                                  const float* __restrict__ roww,
                                                                  float* __restrict__ partial /* [nblk][3][D] */, int64_t T,
                                                                  int D) {
  __shared__ float s_a[kBnThreads], s_b[kBnThreads], s_n[kBnThreads];
  const int slots = kBnThreads / D;
  const int c = threadIdx.x % D, rs = threadIdx.x / D;
  const int64_t rows_per = (T + gridDim.x - 1) / gridDim.x;
  const int64_t lo = (int64_t)blockIdx.x * rows_per, hi = lo + rows_per < T ? lo + rows_per : T;
  float sa = 0.f, sb = 0.f, sn = 0.f;
  if (rs < slots) {
    const float mu = mean[c], rs_ = rstd[c];
    for (int64_t r = lo + rs; r < hi; r += slots) {
      const float w = roww ? roww[r] : 1.0f;
      if (w != 0.0f) {
        const float g = dy[r * D + c];
        sa += g;
        sb = fmaf(g, (z[r * D + c] - mu) * rs_, sb);
        sn += 1.0f;
      }
    }
  }
  s_a[threadIdx.x] = sa, s_b[threadIdx.x] = sb, s_n[threadIdx.x] = sn;
  __syncthreads();
  if (threadIdx.x < D) {
    float A = 0.f, Bv = 0.f, N = 0.f;
    for (int s = 0; s < slots; ++s) A += s_a[s * D + c], Bv += s_b[s * D + c], N += s_n[s * D + c];
    float* p = partial + (size_t)blockIdx.x * 3 * D;
    p[c] = A, p[D + c] = Bv, p[2 * D + c] = N;
  }
}

__global__ void __launch_bounds__(kBnThreads) bn_bwd_apply_kernel(
    const float* __restrict__ dy, const float* __restrict__ z, const float* __restrict__ mean,
    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ bscale,
    const float* __restrict__ roww, const float* __restrict__ partial, int nblk, float* __restrict__ dz,
    float* __restrict__ dbs, float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t T, int D) {
  __shared__ float s_k[kBnMaxD], s_ma[kBnMaxD], s_mb[kBnMaxD], s_mu[kBnMaxD], s_rs[kBnMaxD];
  if (threadIdx.x < D) {
    const int c = threadIdx.x;
    float A = 0.f, Bv = 0.f, N = 0.f;
    for (int k = 0; k < nblk; ++k) {
      const float* p = partial + (size_t)k * 3 * D;
      A += p[c], Bv += p[D + c], N += p[2 * D + c];
    }
    const float inv = N > 0.f ? 1.0f / N : 0.f;
    s_k[c] = (gamma ? gamma[c] : 1.0f) * rstd[c];
    s_ma[c] = A * inv, s_mb[c] = Bv * inv, s_mu[c] = mean[c], s_rs[c] = rstd[c];
    if (blockIdx.x == 0) {
      if (dgamma) dgamma[c] = Bv;
      if (dbeta) dbeta[c] = A;
    }
  }
  __syncthreads();
  const int64_t total = T * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % D);
    const int64_t r = i / D;
    const float w = roww ? roww[r] : 1.0f;
    float g = 0.0f;
    if (w != 0.0f) {
      const float xh = (z[i] - s_mu[c]) * s_rs[c];
      g = s_k[c] * (dy[i] - s_ma[c] - xh * s_mb[c]);
    }
    dz[i] = g;
    if (dbs) dbs[i] = bscale[r] * g;
  }
}

}  // namespace feta

using namespace feta;

extern "C" int feta_add_batchnorm_blocks(int64_t T) { return bn_blocks(T); }

extern "C" int feta_add_batchnorm_fwd(const float* a, const float* b, const float* bscale, const float* roww,
                                      const float* gamma, const float* beta, float* y, float* z, float* mean, float* rstd,
                                      float* running_mean, float* running_var, int64_t* num_batches, float* partial,
                                      float momentum, float eps, int64_t T, int D, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(a && y && z && mean && rstd && partial, "add_batchnorm_fwd: NULL pointer");
  FETA_REQUIRE(T >= 1 && D >= 4 && D <= kBnMaxD && D % 4 == 0 && kBnThreads % D == 0,
               "add_batchnorm_fwd: D must divide 256 and be a multiple of 4 (got %d)", D);
  FETA_REQUIRE(!bscale || b, "add_batchnorm_fwd: bscale without b");
  const int nblk = bn_blocks(T);
  bn_stats_kernel<<<nblk, kBnThreads, 0, st>>>(a, b, bscale, roww, z, partial, T, D);
  FETA_LAUNCH_CHECK();
  int64_t g = ceil_div(T * D, (int64_t)kBnThreads * 4);
  if (g > 2 * kNumSMs) g = 2 * kNumSMs;
  bn_apply_kernel<<<(unsigned)g, kBnThreads, 0, st>>>(z, partial, nblk, gamma, beta, y, mean, rstd, running_mean,
                                                      running_var, num_batches, momentum, eps, T, D);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}

extern "C" int feta_add_batchnorm_bwd(const float* dy, const float* z, const float* mean, const float* rstd,
                                      const float* gamma, const float* bscale, const float* roww, float* dz, float* dbs,
                                      float* dgamma, float* dbeta, float* partial, int64_t T, int D, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(dy && z && mean && rstd && dz && partial, "add_batchnorm_bwd: NULL pointer");
  FETA_REQUIRE(T >= 1 && D >= 4 && D <= kBnMaxD && D % 4 == 0 && kBnThreads % D == 0, "add_batchnorm_bwd: bad D %d", D);
  FETA_REQUIRE(!dbs || bscale, "add_batchnorm_bwd: dbs without bscale");
  const int nblk = bn_blocks(T);
  bn_bwd_stats_kernel<<<nblk, kBnThreads, 0, st>>>(dy, z, mean, rstd, roww, partial, T, D);
  FETA_LAUNCH_CHECK();
  int64_t g = ceil_div(T * D, (int64_t)kBnThreads);
  if (g > 4 * kNumSMs) g = 4 * kNumSMs;
  bn_bwd_apply_kernel<<<(unsigned)g, kBnThreads, 0, st>>>(dy, z, mean, rstd, gamma, bscale, roww, partial, nblk, dz, dbs,
                                                          dgamma, dbeta, T, D);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}
