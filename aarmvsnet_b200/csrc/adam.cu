// Adam step (train.py:180,307: optim.Adam(model.parameters(), lr), optimizer.step()) over ONE flat fp32 buffer holding all 90
// parameters (187 203 floats): a single launch instead of torch's per-tensor loop.  Arithmetic follows torch.optim.Adam's
// single-tensor path op for op (no amsgrad, no decoupled weight decay):
//   g = grad + wd * p;  m = lerp(m, g, 1 - b1);  v = b2 * v + (1 - b2) * g * g;
//   denom = sqrt(v) / sqrt(1 - b2^t) + eps;  p -= (lr / (1 - b1^t)) * m / denom.
#include "common.cuh"
#include "../../include/aarmvs.h"
#include <math.h>

namespace {
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, size_t n, float b1, float b2, float eps, float wd,
                                                   float step_size, float bc2_sqrt) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float gi = g[i];
        const float pi = p[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        const float mi = fmaf(1.f - b1, gi - m[i], m[i]);                       // lerp_(grad, 1 - beta1)
        const float vi = fmaf(1.f - b2, gi * gi, b2 * v[i]);                    // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - step_size * (mi / denom);                                   // addcdiv_(exp_avg, denom, value=-step_size)
    }
}
}  // namespace

extern "C" int aarmvs_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int step, void* stream) {
    AARMVS_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1);
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const size_t blocks = (n + 255) / 256;
    adam_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, (cudaStream_t)stream>>>(
        param, grad, exp_avg, exp_avg_sq, n, beta1, beta2, eps, weight_decay, (float)((double)lr / bc1), (float)sqrt(bc2));
    AARMVS_LAUNCH_RESULT();
}
