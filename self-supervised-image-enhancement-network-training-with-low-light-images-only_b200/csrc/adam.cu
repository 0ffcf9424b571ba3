// torch.optim.Adam (defaults: betas (0.9, 0.999), eps 1e-8, no weight decay / amsgrad) as the reference uses it
// (model.py:213, 316), fused over the ONE flat parameter buffer: a single launch instead of 46 x foreach kernels.
// Algorithmic traffic: read p, g, m, v + write p, m, v = 7 x 4 bytes per parameter.
#include "common.cuh"
#include "kernels.h"

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                                   float b1, float b2, float eps, float bc1, float bc2_sqrt,
                                                   float gscale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * gscale;
  const float mi = b1 * m[i] + (1.f - b1) * gi;          // exp_avg.lerp_(grad, 1-beta1)
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;     // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;        // (sqrt(v) / sqrt(bias_correction2)) + eps
  p[i] -= (lr / bc1) * (mi / denom);                     // step_size = lr / bias_correction1
}

extern "C" int sshslie_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 float lr, float beta1, float beta2, float eps, int step, float grad_scale,
                                 void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n < 1 || step < 1) {
    ss_set_error("sshslie_adam_step: bad argument");
    return SSHSLIE_ERR_ARG;
  }
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale);
  return ss_check_launch("adam_step");
}
