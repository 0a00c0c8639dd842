// Backward pass of the path, part 3: the FeatNet pieces that are not plain convolutions (SURVEY.md §8 rows a13, a14
// backwards; models/drmvsnet.py:19-24, models/module.py:133-189).
//
//   feat_assemble_bwd_kernel   d feature -> d (GroupNorm+ReLU'd 1x1 outputs): a copy for the full-resolution 16 channels, the
//                              transpose of the x2 / x4 bilinear upsampling (align_corners=True) for the other 2 x 8
//   deform_sample_kernel       the deformable GATHER alone: samp[p][k][c] (bilinear sample of tap k) and samp * sigmoid(m);
//                              the 9x32 -> 32 contraction and its two gradients are k = 1 calls of aarmvs_conv_nhwc /
//                              aarmvs_conv_bwd_weight over these [P][288] tensors
//   deform_bwd_kernel          d (samp * m) -> d x (scatter to the four clamped corners, red.global.add.v4.f32), d offsets
//                              (through the bilinear weights g_*, module.py:162-165; floor() is detached, module.py:150; the
//                              position clamp passes gradient inside its range) and d modulation logits
// Sample positions, clamping and weights are the forward's (featnet.cu: deform_conv_kernel), op for op.
#include "common.cuh"
#include "../../include/aarmvs.h"

namespace {

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// bilinear, align_corners=True (F.interpolate at drmvsnet.py:22-23): the forward's coordinates (featnet.cu: up_coords)
__device__ __forceinline__ void up_coords_b(int dst, int in, int out, int& i0, int& i1, float& l0, float& l1) {
    const float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.f;
    const float r = scale * (float)dst;
    i0 = (int)r;
    if (i0 > in - 1) i0 = in - 1;
    i1 = i0 + (i0 < in - 1 ? 1 : 0);
    l1 = r - (float)i0;
    l0 = 1.f - l1;
}

__global__ void __launch_bounds__(256) feat_assemble_bwd_kernel(const float* __restrict__ df, int H, int W, float* __restrict__ dm0,
                                                                float* __restrict__ dm1, float* __restrict__ dm2) {
    const int chunk = threadIdx.x & 7;
    const int n = H * W;
    for (int p = blockIdx.x * 32 + (threadIdx.x >> 3); p < n; p += gridDim.x * 32) {
        const int y = p / W, x = p - y * W;
        const float4 g = ldg4(df + (size_t)p * 32 + chunk * 4);
        if (chunk < 4) {
            st4(dm0 + (size_t)p * 16 + chunk * 4, g);
            continue;
        }
        const bool lvl1 = chunk < 6;
        float* m = lvl1 ? dm1 : dm2;
        const int hs = lvl1 ? H / 2 : H / 4, ws = lvl1 ? W / 2 : W / 4;
        const int cl = (lvl1 ? chunk - 4 : chunk - 6) * 4;
        int y0, y1, x0, x1;
        float ly0, ly1, lx0, lx1;
        up_coords_b(y, hs, H, y0, y1, ly0, ly1);
        up_coords_b(x, ws, W, x0, x1, lx0, lx1);
        const float w00 = ly0 * lx0, w01 = ly0 * lx1, w10 = ly1 * lx0, w11 = ly1 * lx1;
        red_add_v4(m + ((size_t)y0 * ws + x0) * 8 + cl, make_float4(w00 * g.x, w00 * g.y, w00 * g.z, w00 * g.w));
        red_add_v4(m + ((size_t)y0 * ws + x1) * 8 + cl, make_float4(w01 * g.x, w01 * g.y, w01 * g.z, w01 * g.w));
        red_add_v4(m + ((size_t)y1 * ws + x0) * 8 + cl, make_float4(w10 * g.x, w10 * g.y, w10 * g.z, w10 * g.w));
        red_add_v4(m + ((size_t)y1 * ws + x1) * 8 + cl, make_float4(w11 * g.x, w11 * g.y, w11 * g.z, w11 * g.w));
    }
}

// Sample geometry of tap k of pixel (i, j): the forward's arithmetic (featnet.cu), kept in one place for both kernels.
struct DeformTap {
    float g_lt, g_rb, g_lb, g_rt;      // bilinear weights from the clamped corners and position
    float r0, r1, c0, c1, pr, pc;      // clamped corner rows / columns and position (padded coordinates)
    bool pass_r, pass_c;               // the position clamp passes the gradient (0 <= raw <= max)
    long o_lt, o_rb, o_lb, o_rt;       // pixel offsets into the un-padded map, -1 on the zero padding ring
    float m;                           // sigmoid(modulation logit)
};

__device__ __forceinline__ DeformTap deform_tap(const float* om, int i, int j, int k, int H, int W) {
    DeformTap t;
    const float hmax = (float)(H + 1), wmax = (float)(W + 1);      // padded size - 1 (module.py:153-159)
    const float pr_raw = __fadd_rn(__fadd_rn((float)(i + 1), (float)(k / 3 - 1)), om[k]);
    const float pc_raw = __fadd_rn(__fadd_rn((float)(j + 1), (float)(k % 3 - 1)), om[9 + k]);
    t.m = sigmoidf_acc(om[18 + k]);
    const float fr = floorf(pr_raw), fc = floorf(pc_raw);
    t.r0 = fminf(fmaxf(fr, 0.f), hmax); t.r1 = fminf(fmaxf(fr + 1.f, 0.f), hmax);
    t.c0 = fminf(fmaxf(fc, 0.f), wmax); t.c1 = fminf(fmaxf(fc + 1.f, 0.f), wmax);
    t.pr = fminf(fmaxf(pr_raw, 0.f), hmax); t.pc = fminf(fmaxf(pc_raw, 0.f), wmax);
    t.pass_r = pr_raw >= 0.f && pr_raw <= hmax;
    t.pass_c = pc_raw >= 0.f && pc_raw <= wmax;
    t.g_lt = __fmul_rn(__fadd_rn(1.f, __fsub_rn(t.r0, t.pr)), __fadd_rn(1.f, __fsub_rn(t.c0, t.pc)));
    t.g_rb = __fmul_rn(__fsub_rn(1.f, __fsub_rn(t.r1, t.pr)), __fsub_rn(1.f, __fsub_rn(t.c1, t.pc)));
    t.g_lb = __fmul_rn(__fadd_rn(1.f, __fsub_rn(t.r0, t.pr)), __fsub_rn(1.f, __fsub_rn(t.c1, t.pc)));
    t.g_rt = __fmul_rn(__fsub_rn(1.f, __fsub_rn(t.r1, t.pr)), __fadd_rn(1.f, __fsub_rn(t.c0, t.pc)));
    const int ir0 = (int)t.r0 - 1, ir1 = (int)t.r1 - 1, ic0 = (int)t.c0 - 1, ic1 = (int)t.c1 - 1;
    const bool vr0 = ir0 >= 0 && ir0 < H, vr1 = ir1 >= 0 && ir1 < H, vc0 = ic0 >= 0 && ic0 < W, vc1 = ic1 >= 0 && ic1 < W;
    t.o_lt = (vr0 && vc0) ? (long)ir0 * W + ic0 : -1;
    t.o_rb = (vr1 && vc1) ? (long)ir1 * W + ic1 : -1;
    t.o_lb = (vr0 && vc1) ? (long)ir0 * W + ic1 : -1;
    t.o_rt = (vr1 && vc0) ? (long)ir1 * W + ic0 : -1;
    return t;
}

__device__ __forceinline__ float4 corner(const float* x, long o, int sub) {
    return o >= 0 ? ldg4(x + (size_t)o * 32 + sub * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256) deform_sample_kernel(const float* __restrict__ x, int H, int W, const float* __restrict__ offmask,
                                                            float* __restrict__ samp, float* __restrict__ samp_m) {
    const int sub = threadIdx.x & 7;
    const int n = H * W;
    for (int p = blockIdx.x * 32 + (threadIdx.x >> 3); p < n; p += gridDim.x * 32) {
        const int i = p / W, j = p - i * W;
        const float* om = offmask + (size_t)p * 32;
        for (int k = 0; k < 9; ++k) {
            const DeformTap t = deform_tap(om, i, j, k, H, W);
            const float4 lt = corner(x, t.o_lt, sub), rb = corner(x, t.o_rb, sub), lb = corner(x, t.o_lb, sub), rt = corner(x, t.o_rt, sub);
            float4 s;      // module.py:174-177
            s.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.g_lt, lt.x), __fmul_rn(t.g_rb, rb.x)), __fmul_rn(t.g_lb, lb.x)), __fmul_rn(t.g_rt, rt.x));
            s.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.g_lt, lt.y), __fmul_rn(t.g_rb, rb.y)), __fmul_rn(t.g_lb, lb.y)), __fmul_rn(t.g_rt, rt.y));
            s.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.g_lt, lt.z), __fmul_rn(t.g_rb, rb.z)), __fmul_rn(t.g_lb, lb.z)), __fmul_rn(t.g_rt, rt.z));
            s.w = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.g_lt, lt.w), __fmul_rn(t.g_rb, rb.w)), __fmul_rn(t.g_lb, lb.w)), __fmul_rn(t.g_rt, rt.w));
            const size_t o = ((size_t)p * 9 + k) * 32 + sub * 4;
            st4(samp + o, s);
            st4(samp_m + o, make_float4(s.x * t.m, s.y * t.m, s.z * t.m, s.w * t.m));      // module.py:184
        }
    }
}

__device__ __forceinline__ float group8_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }

__global__ void __launch_bounds__(256) deform_bwd_kernel(const float* __restrict__ x, int H, int W, const float* __restrict__ offmask,
                                                         const float* __restrict__ samp, const float* __restrict__ dsamp_m,
                                                         float* __restrict__ dx, float* __restrict__ doffmask) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7;
    const int n = H * W;
    for (int pw = (blockIdx.x * 8 + warp) * 4; pw < n; pw += gridDim.x * 32) {       // a warp's four pixels loop together (shuffles)
        const int p = pw + (lane >> 3);
        const bool valid = p < n;
        const int pc = valid ? p : n - 1;
        const int i = pc / W, j = pc - i * W;
        const float* om = offmask + (size_t)pc * 32;
        for (int k = 0; k < 9; ++k) {
            const DeformTap t = deform_tap(om, i, j, k, H, W);
            const size_t o = ((size_t)pc * 9 + k) * 32 + sub * 4;
            const float4 dsm = ldg4(dsamp_m + o), s = ldg4(samp + o);
            const float dm = group8_sum(dot4(dsm, s));
            const float4 ds = make_float4(dsm.x * t.m, dsm.y * t.m, dsm.z * t.m, dsm.w * t.m);
            const float4 lt = corner(x, t.o_lt, sub), rb = corner(x, t.o_rb, sub), lb = corner(x, t.o_lb, sub), rt = corner(x, t.o_rt, sub);
            const float dg_lt = group8_sum(dot4(ds, lt)), dg_rb = group8_sum(dot4(ds, rb));
            const float dg_lb = group8_sum(dot4(ds, lb)), dg_rt = group8_sum(dot4(ds, rt));
            if (valid) {
                if (t.o_lt >= 0) red_add_v4(dx + (size_t)t.o_lt * 32 + sub * 4, make_float4(t.g_lt * ds.x, t.g_lt * ds.y, t.g_lt * ds.z, t.g_lt * ds.w));
                if (t.o_rb >= 0) red_add_v4(dx + (size_t)t.o_rb * 32 + sub * 4, make_float4(t.g_rb * ds.x, t.g_rb * ds.y, t.g_rb * ds.z, t.g_rb * ds.w));
                if (t.o_lb >= 0) red_add_v4(dx + (size_t)t.o_lb * 32 + sub * 4, make_float4(t.g_lb * ds.x, t.g_lb * ds.y, t.g_lb * ds.z, t.g_lb * ds.w));
                if (t.o_rt >= 0) red_add_v4(dx + (size_t)t.o_rt * 32 + sub * 4, make_float4(t.g_rt * ds.x, t.g_rt * ds.y, t.g_rt * ds.z, t.g_rt * ds.w));
                if (sub == 0) {
                    // factors of the four weights: a = 1 + (r0 - pr), b = 1 - (r1 - pr), c = 1 + (c0 - pc), d = 1 - (c1 - pc)
                    //   g_lt = a c, g_rb = b d, g_lb = a d, g_rt = b c;   da/dpr = -1, db/dpr = +1, dc/dpc = -1, dd/dpc = +1
                    const float fa = 1.f + (t.r0 - t.pr), fb = 1.f - (t.r1 - t.pr), fc = 1.f + (t.c0 - t.pc), fd = 1.f - (t.c1 - t.pc);
                    const float dpr = -fc * dg_lt + fd * dg_rb - fd * dg_lb + fc * dg_rt;
                    const float dpc = -fa * dg_lt + fb * dg_rb + fa * dg_lb - fb * dg_rt;
                    float* q = doffmask + (size_t)p * 32;
                    q[k] = t.pass_r ? dpr : 0.f;
                    q[9 + k] = t.pass_c ? dpc : 0.f;
                    q[18 + k] = dm * t.m * (1.f - t.m);
                }
            }
        }
        if (valid && sub < 5) doffmask[(size_t)p * 32 + 27 + sub] = 0.f;
    }
}

static int blocks_for(size_t groups) {
    const size_t cap = 148 * 16;
    return (int)(groups < cap ? (groups ? groups : 1) : cap);
}

}  // namespace

extern "C" int aarmvs_feat_assemble_bwd(const float* dfeature, int H, int W, float* dm0, float* dm1, float* dm2, void* stream) {
    AARMVS_CHECK_ARG(dfeature && dm0 && dm1 && dm2 && H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0);
    feat_assemble_bwd_kernel<<<blocks_for(((size_t)H * W + 31) / 32), 256, 0, (cudaStream_t)stream>>>(dfeature, H, W, dm0, dm1, dm2);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_deform_sample(const float* x, int H, int W, const float* offmask, float* samp, float* samp_m, void* stream) {
    AARMVS_CHECK_ARG(x && offmask && samp && samp_m && H > 0 && W > 0);
    deform_sample_kernel<<<blocks_for(((size_t)H * W + 31) / 32), 256, 0, (cudaStream_t)stream>>>(x, H, W, offmask, samp, samp_m);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_deform_bwd(const float* x, int H, int W, const float* offmask, const float* samp, const float* dsamp_m, float* dx,
                                 float* doffmask, void* stream) {
    AARMVS_CHECK_ARG(x && offmask && samp && dsamp_m && dx && doffmask && H > 0 && W > 0);
    deform_bwd_kernel<<<blocks_for(((size_t)H * W + 31) / 32), 256, 0, (cudaStream_t)stream>>>(x, H, W, offmask, samp, dsamp_m, dx, doffmask);
    AARMVS_LAUNCH_RESULT();
}
