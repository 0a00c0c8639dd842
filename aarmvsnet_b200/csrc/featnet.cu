// FeatNet pieces that are not plain convolutions:
//   deform_conv_kernel    modulated deformable 3x3 conv (models/module.py:133-189) as a deformable GATHER from the
//                         NHWC input followed by the 9x32 -> 32 contraction; the reference's four [B,32,H,W,9]
//                         gathered tensors and int64 index tensors are never built
//   feat_assemble_kernel  GroupNorm+ReLU of the three 1x1 heads, x2 / x4 bilinear upsampling (align_corners=True)
//                         and the channel concat (models/drmvsnet.py:19-24) -> feature map [H][W][32]
#include "common.cuh"
#include "../../include/aarmvs.h"

struct DeformArgs {
    const float* x;            // layer input, pre-norm NHWC [H][W][32] (GroupNorm+ReLU applied on load) or plain
    const double* x_stats;
    const float* x_gamma;
    const float* x_beta;
    double x_count;
    int x_groups, x_mode;
    int H, W;
    const float* offmask;      // [H][W][32]
    const float* weight;       // [9][32][32]
    const float* bias;
    float* out;
    double* out_stats;         // 4 groups
};

__global__ void __launch_bounds__(128) deform_conv_kernel(const __grid_constant__ DeformArgs a) {
    extern __shared__ __align__(16) float sw[];          // 9*32*32 weights
    __shared__ float s_scale[32], s_shift[32];
    for (int i = threadIdx.x * 4; i < 9 * 1024; i += blockDim.x * 4) st4(sw + i, ldg4(a.weight + i));
    if (threadIdx.x < 32) {
        float sc = 1.f, sh = 0.f;
        if (a.x_mode == AARMVS_SRC_GN_RELU) {
            float mean, rstd;
            gn_mean_rstd(a.x_stats + (threadIdx.x / (32 / a.x_groups)) * 2, a.x_count, mean, rstd);
            sc = rstd * a.x_gamma[threadIdx.x];
            sh = a.x_beta[threadIdx.x] - mean * sc;
        }
        s_scale[threadIdx.x] = sc;
        s_shift[threadIdx.x] = sh;
    }
    __syncthreads();
    const bool norm = a.x_mode == AARMVS_SRC_GN_RELU;
    const int n = a.H * a.W, lane = threadIdx.x & 31;
    const float hmax = (float)(a.H + 1), wmax = (float)(a.W + 1);      // padded size - 1 (module.py:153-159)
    float gs[4] = {0.f, 0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p0 = blockIdx.x * blockDim.x; p0 < n; p0 += gridDim.x * blockDim.x) {
        const int p = p0 + threadIdx.x;
        if (p >= n) continue;
        const int i = p / a.W, j = p - i * a.W;
        const float* om = a.offmask + (size_t)p * 32;
        float acc[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[k] = a.bias[k];
#pragma unroll 1
        for (int k = 0; k < 9; ++k) {
            // p = p_0 + p_n + offset in the zero-padded image (module.py:201-219): centre (i+1, j+1), tap (k/3-1, k%3-1)
            const float pr_raw = __fadd_rn(__fadd_rn((float)(i + 1), (float)(k / 3 - 1)), om[k]);
            const float pc_raw = __fadd_rn(__fadd_rn((float)(j + 1), (float)(k % 3 - 1)), om[9 + k]);
            const float m = sigmoidf_acc(om[18 + k]);
            const float fr = floorf(pr_raw), fc = floorf(pc_raw);
            const float r0 = fminf(fmaxf(fr, 0.f), hmax), r1 = fminf(fmaxf(fr + 1.f, 0.f), hmax);
            const float c0 = fminf(fmaxf(fc, 0.f), wmax), c1 = fminf(fmaxf(fc + 1.f, 0.f), wmax);
            const float pr = fminf(fmaxf(pr_raw, 0.f), hmax), pc = fminf(fmaxf(pc_raw, 0.f), wmax);
            // module.py:162-165 (weights from the CLAMPED corners and position)
            const float g_lt = __fmul_rn(__fadd_rn(1.f, __fsub_rn(r0, pr)), __fadd_rn(1.f, __fsub_rn(c0, pc)));
            const float g_rb = __fmul_rn(__fsub_rn(1.f, __fsub_rn(r1, pr)), __fsub_rn(1.f, __fsub_rn(c1, pc)));
            const float g_lb = __fmul_rn(__fadd_rn(1.f, __fsub_rn(r0, pr)), __fsub_rn(1.f, __fsub_rn(c1, pc)));
            const float g_rt = __fmul_rn(__fsub_rn(1.f, __fsub_rn(r1, pr)), __fadd_rn(1.f, __fsub_rn(c0, pc)));
            // padded coordinates -> unpadded; the padding ring (index 0 and size+1) holds zeros
            const int ir0 = (int)r0 - 1, ir1 = (int)r1 - 1, ic0 = (int)c0 - 1, ic1 = (int)c1 - 1;
            const bool vr0 = ir0 >= 0 && ir0 < a.H, vr1 = ir1 >= 0 && ir1 < a.H;
            const bool vc0 = ic0 >= 0 && ic0 < a.W, vc1 = ic1 >= 0 && ic1 < a.W;
            const float* q_lt = (vr0 && vc0) ? a.x + ((size_t)ir0 * a.W + ic0) * 32 : nullptr;
            const float* q_rb = (vr1 && vc1) ? a.x + ((size_t)ir1 * a.W + ic1) * 32 : nullptr;
            const float* q_lb = (vr0 && vc1) ? a.x + ((size_t)ir0 * a.W + ic1) * 32 : nullptr;
            const float* q_rt = (vr1 && vc0) ? a.x + ((size_t)ir1 * a.W + ic0) * 32 : nullptr;
            const float* wk = sw + k * 1024;
#pragma unroll 2
            for (int c4 = 0; c4 < 8; ++c4) {
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 lt = q_lt ? ldg4(q_lt + c4 * 4) : z, rb = q_rb ? ldg4(q_rb + c4 * 4) : z;
                float4 lb = q_lb ? ldg4(q_lb + c4 * 4) : z, rt = q_rt ? ldg4(q_rt + c4 * 4) : z;
                float v[4];
                const float ltv[4] = {lt.x, lt.y, lt.z, lt.w}, rbv[4] = {rb.x, rb.y, rb.z, rb.w};
                const float lbv[4] = {lb.x, lb.y, lb.z, lb.w}, rtv[4] = {rt.x, rt.y, rt.z, rt.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int c = c4 * 4 + e;
                    float x_lt = ltv[e], x_rb = rbv[e], x_lb = lbv[e], x_rt = rtv[e];
                    if (norm) {
                        x_lt = q_lt ? fmaxf(fmaf(x_lt, s_scale[c], s_shift[c]), 0.f) : 0.f;
                        x_rb = q_rb ? fmaxf(fmaf(x_rb, s_scale[c], s_shift[c]), 0.f) : 0.f;
                        x_lb = q_lb ? fmaxf(fmaf(x_lb, s_scale[c], s_shift[c]), 0.f) : 0.f;
                        x_rt = q_rt ? fmaxf(fmaf(x_rt, s_scale[c], s_shift[c]), 0.f) : 0.f;
                    }
                    // module.py:174-177 then :184
                    float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(g_lt, x_lt), __fmul_rn(g_rb, x_rb)), __fmul_rn(g_lb, x_lb)),
                                        __fmul_rn(g_rt, x_rt));
                    v[e] = __fmul_rn(s, m);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float* wr = wk + (c4 * 4 + e) * 32;
#pragma unroll
                    for (int o = 0; o < 32; o += 4) {
                        const float4 w4 = ld4(wr + o);
                        acc[o] = fmaf(v[e], w4.x, acc[o]);
                        acc[o + 1] = fmaf(v[e], w4.y, acc[o + 1]);
                        acc[o + 2] = fmaf(v[e], w4.z, acc[o + 2]);
                        acc[o + 3] = fmaf(v[e], w4.w, acc[o + 3]);
                    }
                }
            }
        }
        float* o = a.out + (size_t)p * 32;
#pragma unroll
        for (int k = 0; k < 32; k += 4) {
            st4(o + k, make_float4(acc[k], acc[k + 1], acc[k + 2], acc[k + 3]));
            gs[k / 8] += (acc[k] + acc[k + 1]) + (acc[k + 2] + acc[k + 3]);
            gq[k / 8] += (acc[k] * acc[k] + acc[k + 1] * acc[k + 1]) + (acc[k + 2] * acc[k + 2] + acc[k + 3] * acc[k + 3]);
        }
    }
    if (a.out_stats) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            double s = warp_sum((double)gs[g]), q = warp_sum((double)gq[g]);
            if (lane == 0) {
                atomicAdd(a.out_stats + g * 2 + 0, s);
                atomicAdd(a.out_stats + g * 2 + 1, q);
            }
        }
    }
}

extern "C" int aarmvs_deform_conv(const aarmvs_conv_src* x, int H, int W, const float* offmask, const float* weight,
                                  const float* bias, float* out, double* out_gn_sums, void* stream) {
    AARMVS_CHECK_ARG(x && x->ptr && offmask && weight && bias && out && H > 0 && W > 0 && x->channels == 32);
    AARMVS_CHECK_ARG(x->mode == AARMVS_SRC_PLAIN || x->mode == AARMVS_SRC_GN_RELU);
    if (x->mode == AARMVS_SRC_GN_RELU)
        AARMVS_CHECK_ARG(x->gn_sums && x->gn_weight && x->gn_bias && x->gn_groups >= 1 && 32 % x->gn_groups == 0 && x->gn_count > 0);
    DeformArgs a = {};
    a.x = x->ptr; a.x_stats = x->gn_sums; a.x_gamma = x->gn_weight; a.x_beta = x->gn_bias; a.x_count = x->gn_count;
    a.x_groups = x->gn_groups; a.x_mode = x->mode; a.H = H; a.W = W; a.offmask = offmask; a.weight = weight; a.bias = bias;
    a.out = out; a.out_stats = out_gn_sums;
    const int smem = 9 * 1024 * sizeof(float);
    static unsigned long long smem_ok = 0;
    if (cudaError_t e = ensure_dyn_smem(deform_conv_kernel, smem, smem_ok)) return (int)e;
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    int blocks = (H * W + 127) / 128;
    if (blocks > sms * 6) blocks = sms * 6;
    deform_conv_kernel<<<blocks, 128, smem, (cudaStream_t)stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}

// -------------------------------------------------------------------------------------------------------------
struct AssembleArgs {
    const float *m0, *m1, *m2;
    const double *s0, *s1, *s2;
    const float *gw0, *gb0, *gw1, *gb1, *gw2, *gb2;
    int H, W;
    float* out;
};

// bilinear, align_corners=True (F.interpolate at drmvsnet.py:22-23): src = dst * (in - 1) / (out - 1)
__device__ __forceinline__ void up_coords(int dst, int in, int out, int& i0, int& i1, float& l0, float& l1) {
    const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    const float r = scale * (float)dst;
    i0 = (int)r;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = r - (float)i0;
    l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) feat_assemble_kernel(const __grid_constant__ AssembleArgs a) {
    __shared__ float sc[32], sh[32];
    if (threadIdx.x < 32) {
        const int c = threadIdx.x;
        float mean, rstd, g, b;
        if (c < 16) {
            gn_mean_rstd(a.s0 + (c / 8) * 2, 8.0 * a.H * a.W, mean, rstd); g = a.gw0[c]; b = a.gb0[c];
        } else if (c < 24) {
            gn_mean_rstd(a.s1, 8.0 * (a.H / 2) * (a.W / 2), mean, rstd); g = a.gw1[c - 16]; b = a.gb1[c - 16];
        } else {
            gn_mean_rstd(a.s2, 8.0 * (a.H / 4) * (a.W / 4), mean, rstd); g = a.gw2[c - 24]; b = a.gb2[c - 24];
        }
        sc[c] = rstd * g;
        sh[c] = b - mean * rstd * g;
    }
    __syncthreads();
    const int chunk = threadIdx.x & 7;       // 8 lanes x float4 = 32 channels of one pixel
    const int n = a.H * a.W;
    for (int p = blockIdx.x * 32 + (threadIdx.x >> 3); p < n; p += gridDim.x * 32) {
        const int y = p / a.W, x = p - y * a.W;
        float o[4];
        if (chunk < 4) {
            const float4 v = ldg4(a.m0 + (size_t)p * 16 + chunk * 4);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = fmaxf(fmaf(vv[e], sc[chunk * 4 + e], sh[chunk * 4 + e]), 0.f);
        } else {
            const bool lvl1 = chunk < 6;
            const float* m = lvl1 ? a.m1 : a.m2;
            const int hs = lvl1 ? a.H / 2 : a.H / 4, ws = lvl1 ? a.W / 2 : a.W / 4;
            const int cb = lvl1 ? 16 : 24, cl = (lvl1 ? chunk - 4 : chunk - 6) * 4;
            int y0, y1, x0, x1;
            float ly0, ly1, lx0, lx1;
            up_coords(y, hs, a.H, y0, y1, ly0, ly1);
            up_coords(x, ws, a.W, x0, x1, lx0, lx1);
            const float4 v00 = ldg4(m + ((size_t)y0 * ws + x0) * 8 + cl), v01 = ldg4(m + ((size_t)y0 * ws + x1) * 8 + cl);
            const float4 v10 = ldg4(m + ((size_t)y1 * ws + x0) * 8 + cl), v11 = ldg4(m + ((size_t)y1 * ws + x1) * 8 + cl);
            const float a00[4] = {v00.x, v00.y, v00.z, v00.w}, a01[4] = {v01.x, v01.y, v01.z, v01.w};
            const float a10[4] = {v10.x, v10.y, v10.z, v10.w}, a11[4] = {v11.x, v11.y, v11.z, v11.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float s = sc[cb + cl + e], t = sh[cb + cl + e];
                const float n00 = fmaxf(fmaf(a00[e], s, t), 0.f), n01 = fmaxf(fmaf(a01[e], s, t), 0.f);
                const float n10 = fmaxf(fmaf(a10[e], s, t), 0.f), n11 = fmaxf(fmaf(a11[e], s, t), 0.f);
                o[e] = ly0 * (lx0 * n00 + lx1 * n01) + ly1 * (lx0 * n10 + lx1 * n11);
            }
        }
        st4(a.out + (size_t)p * 32 + chunk * 4, make_float4(o[0], o[1], o[2], o[3]));
    }
}

extern "C" int aarmvs_feat_assemble(const float* m0, const float* m1, const float* m2, const double* sums0,
                                    const double* sums1, const double* sums2, const float* gnw0, const float* gnb0,
                                    const float* gnw1, const float* gnb1, const float* gnw2, const float* gnb2, int H,
                                    int W, float* feature, void* stream) {
    AARMVS_CHECK_ARG(m0 && m1 && m2 && sums0 && sums1 && sums2 && gnw0 && gnb0 && gnw1 && gnb1 && gnw2 && gnb2 && feature);
    AARMVS_CHECK_ARG(H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0);
    AssembleArgs a = {m0, m1, m2, sums0, sums1, sums2, gnw0, gnb0, gnw1, gnb1, gnw2, gnb2, H, W, feature};
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    int blocks = (H * W + 31) / 32;
    if (blocks > sms * 8) blocks = sms * 8;
    feat_assemble_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}
