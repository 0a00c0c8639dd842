// Backward pass of the path, part 2: the plane sweep (SURVEY.md §8 rows a3-a5, a7 backwards; models/module.py:6-38,
// models/drmvsnet.py:27-38,311-319).  For one depth plane, with c_v = (warp_v - ref)^2 and cost = -(1/S) sum_v (1 + w_v) c_v:
//
//   aggregate_bwd_kernel   d cost -> d c_v (direct term) and d w_v
//   omega_bwd_kernel<3..0> d w_v back through sigmoid, the 1x1 convs and the three GroupNorms of omega's tail down to the
//                          stem conv output t0.  Every GroupNorm backward needs two sums over the whole map, so this is four
//                          streaming passes over t0 (16 B per pixel and view), each re-evaluating the point-wise chain --
//                          the mirror image of the forward's passes A-D.  Parameter gradients are reduced per block and
//                          added to an fp64 accumulator in OmegaTail order.
//   (the stem conv's d c_v += d t0 (*) W^T and dW are aarmvs_conv_nhwc / aarmvs_conv_bwd_weight calls)
//   warp_bwd_kernel        d c_v -> d ref and, through the four bilinear taps of the forward's sample positions (same
//                          bit-exact homography, common.cuh), d src as vector reductions red.global.add.v4.f32.  The grid is
//                          built under torch.no_grad() in the reference (module.py:15-34): cameras and depths get no gradient.
#include "common.cuh"
#include "omega.cuh"
#include "../../include/aarmvs.h"
#include <string.h>

namespace {

__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ------------------------------------------------------------------------------------------------ aggregation
__global__ void __launch_bounds__(256) aggregate_bwd_kernel(const float* __restrict__ dcost, const float* __restrict__ costv,
                                                            const float* __restrict__ reweight, int S, size_t P,
                                                            float* __restrict__ dcostv, float* __restrict__ dreweight) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 7;
    const float inv_s = 1.0f / (float)S;
    for (size_t pw = ((size_t)blockIdx.x * 8 + warp) * 4; pw < P; pw += (size_t)gridDim.x * 32) {
        const size_t p = pw + (lane >> 3);
        const bool valid = p < P;
        const size_t pc = valid ? p : P - 1;
        const float4 dc = ldg4(dcost + pc * 32 + sub * 4);
        for (int v = 0; v < S; ++v) {
            const float4 c = ldg4(costv + ((size_t)v * P + pc) * 32 + sub * 4);
            float dot = (c.x * dc.x + c.y * dc.y) + (c.z * dc.z + c.w * dc.w);
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            dot += __shfl_xor_sync(0xffffffffu, dot, 4);
            if (valid) {
                const float f = -(1.0f + __ldg(reweight + (size_t)v * P + p)) * inv_s;
                st4(dcostv + ((size_t)v * P + p) * 32 + sub * 4, make_float4(f * dc.x, f * dc.y, f * dc.z, f * dc.w));
                if (sub == 0) dreweight[(size_t)v * P + p] = -dot * inv_s;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ omega tail
struct OmegaBwdArgs {
    const float* t0;           // [S][P][4]
    const float* dw;           // [S][P]
    size_t P;
    OmegaTail q;
    const double* stats;       // [S][3][2] forward GroupNorm sums
    double* bstats;            // [S][3][2] backward sums: slot k = GroupNorm k+1: (sum dxhat * xhat, sum dxhat)
    double* dtail;             // [69], OmegaTail order
    float* dt0;                // [S][P][4]
};

// offsets into dtail (floats of OmegaTail)
constexpr int T_G1 = 0, T_B1 = 4, T_W1 = 8, T_C1 = 24, T_G2 = 28, T_B2 = 32, T_W2 = 36, T_C2 = 52, T_G3 = 56, T_B3 = 60, T_W3 = 64,
              T_C3 = 68;
static_assert(sizeof(OmegaTail) == 69 * sizeof(float), "OmegaTail layout");

template <int STAGE>
__global__ void __launch_bounds__(256) omega_bwd_kernel(const __grid_constant__ OmegaBwdArgs a) {
    constexpr int NV = STAGE == 3 ? 15 : STAGE == 0 ? 1 : 30;
    __shared__ double red[NV * 8];
    const OmegaTail& q = a.q;
    const int v = blockIdx.y;
    const double cnt = 4.0 * (double)a.P;
    OmegaNorm n;
    gn_mean_rstd(a.stats + v * 6 + 0, cnt, n.m1, n.r1);
    gn_mean_rstd(a.stats + v * 6 + 2, cnt, n.m2, n.r2);
    gn_mean_rstd(a.stats + v * 6 + 4, cnt, n.m3, n.r3);
    const double* bs = a.bstats + v * 6;
    const float A3 = STAGE < 3 ? (float)(bs[4] / cnt) : 0.f, B3 = STAGE < 3 ? (float)(bs[5] / cnt) : 0.f;
    const float A2 = STAGE < 2 ? (float)(bs[2] / cnt) : 0.f, B2 = STAGE < 2 ? (float)(bs[3] / cnt) : 0.f;
    const float A1 = STAGE < 1 ? (float)(bs[0] / cnt) : 0.f, B1 = STAGE < 1 ? (float)(bs[1] / cnt) : 0.f;
    const float* t0 = a.t0 + (size_t)v * a.P * 4;
    const float* dw = a.dw + (size_t)v * a.P;
    float acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = 0.f;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < a.P; p += (size_t)gridDim.x * blockDim.x) {
        const float4 t4 = ldg4(t0 + p * 4);
        const float t[4] = {t4.x, t4.y, t4.z, t4.w};
        // ---- forward chain (omega.cuh, spelled out because the backward needs every intermediate)
        float x1h[4], y1[4], a1[4], u[4], x2h[4], y2[4], a2[4], vv[4], x3h[4], pre[4], rr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            x1h[k] = (t[k] - n.m1) * n.r1;
            y1[k] = x1h[k] * q.g1[k] + q.b1[k];
            a1[k] = fmaxf(y1[k], 0.f);
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float s = q.c1[o];
#pragma unroll
            for (int i = 0; i < 4; ++i) s = fmaf(q.w1[o * 4 + i], a1[i], s);
            u[o] = s;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            x2h[k] = (u[k] - n.m2) * n.r2;
            y2[k] = x2h[k] * q.g2[k] + q.b2[k];
            a2[k] = fmaxf(y2[k], 0.f);
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float s = q.c2[o];
#pragma unroll
            for (int i = 0; i < 4; ++i) s = fmaf(q.w2[o * 4 + i], a2[i], s);
            vv[o] = s;
        }
        float z = q.c3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            x3h[k] = (vv[k] - n.m3) * n.r3;
            pre[k] = (x3h[k] * q.g3[k] + q.b3[k]) + a1[k];
            rr[k] = fmaxf(pre[k], 0.f);
            z = fmaf(q.w3[k], rr[k], z);
        }
        const float om = sigmoidf_acc(z);
        // ---- backward
        const float dz = __ldg(dw + p) * om * (1.f - om);
        float dpre[4], dxh3[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dpre[k] = pre[k] > 0.f ? dz * q.w3[k] : 0.f;
            dxh3[k] = dpre[k] * q.g3[k];
        }
        if (STAGE == 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc[0] += dxh3[k] * x3h[k];
                acc[1] += dxh3[k];
                acc[2 + k] += dpre[k] * x3h[k];      // d g3
                acc[6 + k] += dpre[k];               // d b3
                acc[10 + k] += dz * rr[k];           // d w3
            }
            acc[14] += dz;                           // d c3
            continue;
        }
        float dvv[4], da2[4], dy2[4], dxh2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) dvv[k] = n.r3 * (dxh3[k] - (x3h[k] * A3 + B3));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float s = 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) s = fmaf(q.w2[o * 4 + i], dvv[o], s);
            da2[i] = s;
            dy2[i] = y2[i] > 0.f ? s : 0.f;
            dxh2[i] = dy2[i] * q.g2[i];
        }
        if (STAGE == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc[0] += dxh2[k] * x2h[k];
                acc[1] += dxh2[k];
                acc[2 + k] += dy2[k] * x2h[k];       // d g2
                acc[6 + k] += dy2[k];                // d b2
                acc[26 + k] += dvv[k];               // d c2
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[10 + k * 4 + i] += dvv[k] * a2[i];      // d w2[o = k][i]
            }
            continue;
        }
        float du[4], dy1[4], dxh1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) du[k] = n.r2 * (dxh2[k] - (x2h[k] * A2 + B2));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float s = dpre[i];                       // the residual connection (module.py:262: out += x)
#pragma unroll
            for (int o = 0; o < 4; ++o) s = fmaf(q.w1[o * 4 + i], du[o], s);
            dy1[i] = y1[i] > 0.f ? s : 0.f;
            dxh1[i] = dy1[i] * q.g1[i];
        }
        if (STAGE == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                acc[0] += dxh1[k] * x1h[k];
                acc[1] += dxh1[k];
                acc[2 + k] += dy1[k] * x1h[k];       // d g1
                acc[6 + k] += dy1[k];                // d b1
                acc[26 + k] += du[k];                // d c1
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[10 + k * 4 + i] += du[k] * a1[i];       // d w1[o = k][i]
            }
            continue;
        }
        st4(a.dt0 + ((size_t)v * a.P + p) * 4,
            make_float4(n.r1 * (dxh1[0] - (x1h[0] * A1 + B1)), n.r1 * (dxh1[1] - (x1h[1] * A1 + B1)),
                        n.r1 * (dxh1[2] - (x1h[2] * A1 + B1)), n.r1 * (dxh1[3] - (x1h[3] * A1 + B1))));
    }
    if (STAGE == 0) return;
    double dv[NV], tot[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) dv[i] = (double)acc[i];
    block_sum<NV, 8>(dv, red, tot);
    if (threadIdx.x == 0) {
        double* b = a.bstats + v * 6 + (STAGE - 1) * 2;
        atomicAdd(b + 0, tot[0]);
        atomicAdd(b + 1, tot[1]);
        const int og = STAGE == 3 ? T_G3 : STAGE == 2 ? T_G2 : T_G1, ob = STAGE == 3 ? T_B3 : STAGE == 2 ? T_B2 : T_B1;
        for (int k = 0; k < 4; ++k) {
            atomicAdd(a.dtail + og + k, tot[2 + k]);
            atomicAdd(a.dtail + ob + k, tot[6 + k]);
        }
        if (STAGE == 3) {
            for (int k = 0; k < 4; ++k) atomicAdd(a.dtail + T_W3 + k, tot[10 + k]);
            atomicAdd(a.dtail + T_C3, tot[14]);
        } else {
            const int ow = STAGE == 2 ? T_W2 : T_W1, oc = STAGE == 2 ? T_C2 : T_C1;
            for (int k = 0; k < 16; ++k) atomicAdd(a.dtail + ow + k, tot[10 + k]);
            for (int k = 0; k < 4; ++k) atomicAdd(a.dtail + oc + k, tot[26 + k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ warp
__global__ void __launch_bounds__(256) warp_bwd_kernel(const float* __restrict__ ref, const float* __restrict__ src, size_t view_stride,
                                                       int H, int W, int S, const float* __restrict__ cams,
                                                       const float* __restrict__ depth_values, int d_host, const int* __restrict__ d_dev,
                                                       int align_corners, const float* __restrict__ dcostv, float* __restrict__ dref,
                                                       float* __restrict__ dsrc) {
    __shared__ float scam[AARMVS_MAX_SRC * 12];
    for (int i = threadIdx.x; i < S * 12; i += blockDim.x) scam[i] = cams[i];
    __syncthreads();
    const float depth = depth_values[plane_index(d_host, d_dev)];
    const size_t P = (size_t)H * W;
    const int sub = threadIdx.x & 7;
    for (size_t p = (size_t)blockIdx.x * 32 + (threadIdx.x >> 3); p < P; p += (size_t)gridDim.x * 32) {
        const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
        const float4 rv = ldg4(ref + p * 32 + sub * 4);
        float4 dr = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int v = 0; v < S; ++v) {
            float ix, iy;
            sweep_position(scam + v * 12, (float)x, (float)y, depth, W, H, align_corners != 0, ix, iy);
            const Taps t = make_taps(ix, iy, W, H);
            const float4 wv = gather_taps(src + (size_t)v * view_stride, t, sub);
            const float4 dc = ldg4(dcostv + ((size_t)v * P + p) * 32 + sub * 4);
            const float4 g = make_float4(2.f * (wv.x - rv.x) * dc.x, 2.f * (wv.y - rv.y) * dc.y, 2.f * (wv.z - rv.z) * dc.z,
                                         2.f * (wv.w - rv.w) * dc.w);
            dr.x -= g.x; dr.y -= g.y; dr.z -= g.z; dr.w -= g.w;
            float* dv = dsrc + (size_t)v * view_stride + sub * 4;
            if (t.o_nw >= 0) red_add_v4(dv + (size_t)t.o_nw * 32, make_float4(t.w_nw * g.x, t.w_nw * g.y, t.w_nw * g.z, t.w_nw * g.w));
            if (t.o_ne >= 0) red_add_v4(dv + (size_t)t.o_ne * 32, make_float4(t.w_ne * g.x, t.w_ne * g.y, t.w_ne * g.z, t.w_ne * g.w));
            if (t.o_sw >= 0) red_add_v4(dv + (size_t)t.o_sw * 32, make_float4(t.w_sw * g.x, t.w_sw * g.y, t.w_sw * g.z, t.w_sw * g.w));
            if (t.o_se >= 0) red_add_v4(dv + (size_t)t.o_se * 32, make_float4(t.w_se * g.x, t.w_se * g.y, t.w_se * g.z, t.w_se * g.w));
        }
        float* q = dref + p * 32 + sub * 4;           // one owner per (pixel, chunk): plain read-modify-write
        const float4 old = ld4(q);
        st4(q, make_float4(old.x + dr.x, old.y + dr.y, old.z + dr.z, old.w + dr.w));
    }
}

}  // namespace

extern "C" int aarmvs_aggregate_bwd(const float* dcost, const float* costv, const float* reweight, int S, size_t P, float* dcostv,
                                    float* dreweight, void* stream) {
    AARMVS_CHECK_ARG(dcost && costv && reweight && dcostv && dreweight && P > 0 && S >= 1 && S <= AARMVS_MAX_SRC);
    const size_t blocks = (P + 31) / 32;
    aggregate_bwd_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(dcost, costv, reweight, S, P,
                                                                                                      dcostv, dreweight);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_omega_bwd(int stage, const float* t0, const float* dreweight, int H, int W, int S, const float* tail_host,
                                const double* stats, double* bstats, double* dtail, float* dt0, void* stream) {
    AARMVS_CHECK_ARG(t0 && dreweight && tail_host && stats && bstats && dtail && H > 0 && W > 0 && S >= 1 && S <= AARMVS_MAX_SRC);
    AARMVS_CHECK_ARG(stage >= 0 && stage <= 3 && (stage != 0 || dt0));
    OmegaBwdArgs a = {};
    a.t0 = t0; a.dw = dreweight; a.P = (size_t)H * W;
    memcpy(&a.q, tail_host, sizeof(OmegaTail));
    a.stats = stats; a.bstats = bstats; a.dtail = dtail; a.dt0 = dt0;
    const size_t blocks = (a.P + 255) / 256;
    dim3 grid((unsigned)(blocks < 148 * 4 ? blocks : 148 * 4), S);
    cudaStream_t st = (cudaStream_t)stream;
    if (stage == 3) omega_bwd_kernel<3><<<grid, 256, 0, st>>>(a);
    else if (stage == 2) omega_bwd_kernel<2><<<grid, 256, 0, st>>>(a);
    else if (stage == 1) omega_bwd_kernel<1><<<grid, 256, 0, st>>>(a);
    else omega_bwd_kernel<0><<<grid, 256, 0, st>>>(a);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_warp_bwd(const float* ref, const float* src, size_t view_stride, int H, int W, int S, const float* cams,
                               const float* depth_values, int d_host, const int* d_dev, int align_corners, const float* dcostv, float* dref,
                               float* dsrc, void* stream) {
    AARMVS_CHECK_ARG(ref && src && cams && depth_values && dcostv && dref && dsrc && H > 0 && W > 0 && S >= 1 && S <= AARMVS_MAX_SRC);
    const size_t blocks = ((size_t)H * W + 31) / 32;
    warp_bwd_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(ref, src, view_stride, H, W, S, cams,
                                                                                                 depth_values, d_host, d_dev, align_corners,
                                                                                                 dcostv, dref, dsrc);
    AARMVS_LAUNCH_RESULT();
}
