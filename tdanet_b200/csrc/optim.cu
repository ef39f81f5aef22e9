// Optimiser step of the reference's training loop on flat parameter / gradient buffers:
// global gradient-norm clipping (Trainer(gradient_clip_val=5.0), audio_train.py:193 ->
// torch.nn.utils.clip_grad_norm_) followed by torch.optim.Adam (configs/tdanet_lsr2.yml:42-45:
// lr 1e-3, betas (0.9, 0.999), eps 1e-8, weight_decay 0), all on the device with no host round trip:
// the norm is read by the update kernel, the step counter lives in device memory so that the whole
// training step can be replayed as one CUDA graph.
#include "kernels.h"

namespace td {

__global__ void sqnorm_kernel(const float* __restrict__ g, size_t n, double* __restrict__ out) {
  grid_dep_wait();
  double acc = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = g[i];
    acc += (double)v * v;
  }
  block_accum2(out, acc, 0.0);
}

struct AdamArgs {
  float* p;
  const float* g;
  float *m, *v;
  size_t n;
  float lr, b1, b2, eps, max_norm, grad_scale;
  const double* sqnorm;  // [2]: sum of squares of the (unscaled) gradient, or null (no clipping)
  const int32_t* step;   // number of steps taken so far
};

__global__ void adam_kernel(AdamArgs a) {
  grid_dep_wait();
  const int t = *a.step + 1;
  // clip_grad_norm_: coef = clamp(max_norm / (total_norm + 1e-6), max = 1)
  float coef = a.grad_scale;
  if (a.sqnorm && a.max_norm > 0.f) {
    const float total = (float)(sqrt(a.sqnorm[0]) * (double)a.grad_scale);
    const float c = a.max_norm / (total + 1e-6f);
    coef *= c < 1.f ? c : 1.f;
  }
  const float bc1 = 1.f - powf(a.b1, (float)t), bc2 = 1.f - powf(a.b2, (float)t);
  const float step_size = a.lr / bc1, rsq_bc2 = 1.f / sqrtf(bc2);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (size_t)gridDim.x * blockDim.x) {
    const float g = a.g[i] * coef;
    const float m = a.m[i] + (g - a.m[i]) * (1.f - a.b1);       // exp_avg.lerp_(grad, 1 - beta1)
    const float v = fmaf(a.b2, a.v[i], (1.f - a.b2) * g * g);   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    a.m[i] = m;
    a.v[i] = v;
    const float denom = sqrtf(v) * rsq_bc2 + a.eps;
    a.p[i] -= step_size * (m / denom);
  }
}

__global__ void step_inc_kernel(int32_t* step) {
  grid_dep_wait();
  *step += 1;
}

}  // namespace td

using namespace td;

extern "C" {

int tdanet_grad_sqnorm(const float* grads, size_t n, double* sqnorm, tdanet_stream_t stream) {
  TD_REQUIRE(grads && sqnorm, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  TD_CUDA(cudaMemsetAsync(sqnorm, 0, 2 * sizeof(double), st));
  size_t blocks = (n + 1023) / 1024;
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  TD_LAUNCH_RED(sqnorm_kernel, (unsigned)blocks, 256, 0, st, grads, n, sqnorm);
  return 0;
}

int tdanet_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n, float lr,
                     float beta1, float beta2, float eps, float max_grad_norm, float grad_scale,
                     const double* sqnorm, int32_t* step, tdanet_stream_t stream) {
  TD_REQUIRE(params && grads && exp_avg && exp_avg_sq && step, "NULL argument");
  TD_REQUIRE(max_grad_norm <= 0.f || sqnorm, "clipping needs the gradient norm (tdanet_grad_sqnorm)");
  cudaStream_t st = (cudaStream_t)stream;
  AdamArgs a{params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, max_grad_norm, grad_scale, sqnorm, step};
  size_t blocks = (n + 1023) / 1024;
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  TD_LAUNCH(adam_kernel, (unsigned)blocks, 256, 0, st, a);
  TD_LAUNCH(step_inc_kernel, 1, 1, 0, st, step);
  return 0;
}

}  // extern "C"
