// Training tail, first step of the backward pass (models/drmvsnet.py:289-291 softmax over the D
// regularised cost maps, :347-381 mvsnet_cls_loss):
//   loss = mean_b [ sum_p mask_p * ( log sum_d exp(cost_d,p) - cost_gt(p),p ) / (sum_p mask_p + 1e-6) ]
//   d loss / d cost_d,p = scale * mask_p * ( softmax_d(cost_:,p) - [d == gt(p)] ),   scale = 1 / ((sum_p mask_p + 1e-6) * B)
// One thread per pixel walks the D planes twice (running max and exp-sum with the max subtracted, as F.softmax does; then the
// gradient): plane-major [D][P] storage makes every access of a warp one coalesced line.  HBM-bound: 3 * 4 * D * P bytes.
#include "common.cuh"
#include "../../include/aarmvs.h"

namespace {

__global__ void __launch_bounds__(256) loss_grad_kernel(const float* __restrict__ cost, const int* __restrict__ gt_index,
                                                        const float* __restrict__ mask, float scale, int D, size_t P,
                                                        float* __restrict__ dcost) {
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
        float mx = -INFINITY;
        for (int d = 0; d < D; ++d) mx = fmaxf(mx, __ldg(cost + (size_t)d * P + p));
        float sum = 0.f;
        for (int d = 0; d < D; ++d) sum += expf(__ldg(cost + (size_t)d * P + p) - mx);
        const float inv = 1.0f / sum, w = scale * mask[p];
        const int gt = gt_index[p];
        for (int d = 0; d < D; ++d) {
            const float prob = expf(__ldg(cost + (size_t)d * P + p) - mx) * inv;
            dcost[(size_t)d * P + p] = w * (prob - (d == gt ? 1.0f : 0.0f));
        }
    }
}

}  // namespace

extern "C" int aarmvs_loss_grad(const float* cost, const int* gt_index, const float* mask, float scale, int D, size_t P,
                                float* dcost, void* stream) {
    AARMVS_CHECK_ARG(cost && gt_index && mask && dcost && D >= 1 && P >= 1);
    const size_t blocks = (P + 255) / 256;
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    const unsigned grid = (unsigned)(blocks < (size_t)sms * 8 ? blocks : (size_t)sms * 8);
    loss_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cost, gt_index, mask, scale, D, P, dcost);
    AARMVS_LAUNCH_RESULT();
}
