// Optimizer step of the measured training step: Adam / AdamW over ONE flat parameter buffer.
//
// The reference drivers step `optim.Adam(model.parameters(), lr=args.lr)`
// (experiments/run_transformer_gengcn.py:302) or `optim.AdamW(..., weight_decay=...)`
// (experiments/run_transformer_gengcn_SBM_cv.py:371).  PyTorch's fused multi-tensor Adam is 4 launches
// (~70 us at the end of every step's critical chain for this model's ~170 small tensors); with parameters,
// gradients and both moments each living in one flat fp32 buffer the whole update is one elementwise kernel.
// The step counter and the learning rate live on the device, so a captured CUDA graph replays correctly and a
// scheduler can change the rate without a re-capture.
#include "common.cuh"

namespace feta {

__global__ void adam_tick_kernel(float* __restrict__ step) { *step += 1.0f; }

struct AdamCoef {
  float step_size, inv_bc2_sqrt, decay;
};
__device__ __forceinline__ AdamCoef adam_coef(const float* step, const float* lr, float b1, float b2, float wd) {
  const float t = *step, l = *lr;
  AdamCoef c;
  c.step_size = l / (1.0f - powf(b1, t));
  c.inv_bc2_sqrt = rsqrtf(1.0f - powf(b2, t));
  c.decay = 1.0f - l * wd;   // decoupled (AdamW); wd = 0: plain Adam
  return c;
}
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamCoef& c, float b1, float b2,
                                         float eps) {
  m = fmaf(b1, m, (1.0f - b1) * g);                  // exp_avg
  v = fmaf(b2, v, (1.0f - b2) * g * g);              // exp_avg_sq
  const float denom = sqrtf(v) * c.inv_bc2_sqrt + eps;
  p = fmaf(-c.step_size, m / denom, p * c.decay);
}

__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                       const float* __restrict__ lr, float b1, float b2, float eps,
                                                       float wd, float gscale, const float* __restrict__ step) {
  const AdamCoef c = adam_coef(step, lr, b1, b2, wd);
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i],
           vv = reinterpret_cast<float4*>(v)[i];
    const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + i);
    adam_one(pp.x, gg.x * gscale, mm.x, vv.x, c, b1, b2, eps);
    adam_one(pp.y, gg.y * gscale, mm.y, vv.y, c, b1, b2, eps);
    adam_one(pp.z, gg.z * gscale, mm.z, vv.z, c, b1, b2, eps);
    adam_one(pp.w, gg.w * gscale, mm.w, vv.w, c, b1, b2, eps);
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    adam_one(p[i], g[i] * gscale, m[i], v[i], c, b1, b2, eps);
}

}  // namespace feta

using namespace feta;

extern "C" int feta_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                              const float* lr, float beta1, float beta2, float eps, float weight_decay,
                              float grad_scale, float* step, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  FETA_REQUIRE(n >= 0 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f,
               "adam_step: bad hyper-parameters");
  if (n == 0) return FETA_OK;
  FETA_REQUIRE(params && grads && exp_avg && exp_avg_sq && lr && step, "adam_step: NULL pointer argument");
  FETA_REQUIRE((((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16) == 0,
               "adam_step: flat buffers must be 16-byte aligned");
  adam_tick_kernel<<<1, 1, 0, st>>>(step);
  FETA_LAUNCH_CHECK();
  int64_t blocks = ceil_div(n / 4 + 1, 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  adam_flat_kernel<<<(unsigned)blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                    weight_decay, grad_scale, step);
  FETA_LAUNCH_CHECK();
  return FETA_OK;
}
