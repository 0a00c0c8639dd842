// The point-wise tail of omega (InterViewAAModule, models/drmvsnet.py:27-38 with module.py:252-264): parameters and the
// forward pieces shared by the plane-sweep passes (warp_aggregate.cu) and their backward (backward_sweep.cu).
#pragma once
#include "common.cuh"

struct OmegaTail {             // the 1x1 layers and GroupNorm affines of omega (models/drmvsnet.py:31-34)
    float g1[4], b1[4];        // reweight_network.0.1   GroupNorm(1, 4)
    float w1[16], c1[4];       // reweight_network.1.stem.0.0  conv1x1 4->4  (w[o*4+i])
    float g2[4], b2[4];        // reweight_network.1.stem.0.1
    float w2[16], c2[4];       // reweight_network.1.stem.1
    float g3[4], b3[4];        // reweight_network.1.stem.2
    float w3[4], c3;           // reweight_network.2  conv1x1 4->1
};

struct OmegaNorm { float m1, r1, m2, r2, m3, r3; };

// a1 = ReLU(GN1(t0));  u = conv1x1(a1)
__device__ __forceinline__ void omega_u(const OmegaTail& q, const OmegaNorm& n, float4 t, float (&a1)[4], float (&u)[4]) {
    const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) a1[k] = fmaxf((tv[k] - n.m1) * n.r1 * q.g1[k] + q.b1[k], 0.f);
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float s = q.c1[o];
#pragma unroll
        for (int i = 0; i < 4; ++i) s = fmaf(q.w1[o * 4 + i], a1[i], s);
        u[o] = s;
    }
}
// a2 = ReLU(GN2(u));  v = conv1x1(a2)
__device__ __forceinline__ void omega_v(const OmegaTail& q, const OmegaNorm& n, const float (&u)[4], float (&vv)[4]) {
    float a2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) a2[k] = fmaxf((u[k] - n.m2) * n.r2 * q.g2[k] + q.b2[k], 0.f);
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        float s = q.c2[o];
#pragma unroll
        for (int i = 0; i < 4; ++i) s = fmaf(q.w2[o * 4 + i], a2[i], s);
        vv[o] = s;
    }
}
// w = sigmoid(conv1x1(ReLU(GN3(v) + a1)))
__device__ __forceinline__ float omega_w(const OmegaTail& q, const OmegaNorm& n, const float (&a1)[4], const float (&vv)[4]) {
    float s = q.c3;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float r = fmaxf(((vv[k] - n.m3) * n.r3 * q.g3[k] + q.b3[k]) + a1[k], 0.f);
        s = fmaf(q.w3[k], r, s);
    }
    return sigmoidf_acc(s);
}
