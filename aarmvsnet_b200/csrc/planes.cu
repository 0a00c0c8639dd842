// fp32 NHWC -> fp16 (hi, lo) operand planes [2][C/8][H][W][8] for the TMA-fed tensor-core kernels (convlstm_tma.cu),
// optionally applying the producer's GroupNorm + ReLU (models/module.py:284-286: deConvGnReLU's norm needs the sums of
// the whole map, so it cannot be applied by the deconvolution kernel itself).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/aarmvs.h"

namespace {

struct SplitArgs {
    const float* src;
    const double* stats;
    const float* gamma;
    const float* beta;
    int ch, mode, groups;
    double count;
    size_t P;
    __half* planes;
};

__global__ void __launch_bounds__(256) split_planes_kernel(const SplitArgs a) {
    __shared__ float s_scale[64], s_shift[64];
    if (threadIdx.x < a.ch) {
        float sc = 1.f, sh = 0.f;
        if (a.mode == AARMVS_SRC_GN_RELU) {
            float mean, rstd;
            gn_mean_rstd(a.stats + (threadIdx.x / (a.ch / a.groups)) * 2, a.count, mean, rstd);
            sc = rstd * a.gamma[threadIdx.x];
            sh = a.beta[threadIdx.x] - mean * sc;
        }
        s_scale[threadIdx.x] = sc;
        s_shift[threadIdx.x] = sh;
    }
    __syncthreads();
    const int q = a.ch >> 2, chk = a.ch >> 3;                  // float4 groups / 16-byte chunks per pixel
    const size_t total = a.P * q, stride = (size_t)gridDim.x * blockDim.x;
    const bool gn = a.mode == AARMVS_SRC_GN_RELU;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t px = i / q;
        const int c4 = (int)(i - px * q);
        float4 v = ldg4(a.src + i * 4);
        if (gn) {
            const float* sc = s_scale + c4 * 4;
            const float* sh = s_shift + c4 * 4;
            v.x = fmaxf(fmaf(v.x, sc[0], sh[0]), 0.f); v.y = fmaxf(fmaf(v.y, sc[1], sh[1]), 0.f);
            v.z = fmaxf(fmaf(v.z, sc[2], sh[2]), 0.f); v.w = fmaxf(fmaf(v.w, sc[3], sh[3]), 0.f);
        }
        uint2 hi, lo;
        tc::split4_f16(v, hi, lo);
        const size_t off = ((size_t)(c4 >> 1) * a.P + px) * 8 + (c4 & 1) * 4;
        *reinterpret_cast<uint2*>(a.planes + off) = hi;
        *reinterpret_cast<uint2*>(a.planes + (size_t)chk * a.P * 8 + off) = lo;
    }
}

}  // namespace

extern "C" int aarmvs_split_planes(const aarmvs_conv_src* src, int H, int W, void* planes, void* stream) {
    AARMVS_CHECK_ARG(src && src->ptr && planes && H > 0 && W > 0);
    AARMVS_CHECK_ARG(src->channels >= 8 && src->channels <= 64 && src->channels % 8 == 0 && src->pp_stride == 0);
    AARMVS_CHECK_ARG(src->mode == AARMVS_SRC_PLAIN || src->mode == AARMVS_SRC_GN_RELU);
    if (src->mode == AARMVS_SRC_GN_RELU)
        AARMVS_CHECK_ARG(src->gn_sums && src->gn_weight && src->gn_bias && src->gn_groups >= 1 && src->gn_count > 0 &&
                         src->channels % src->gn_groups == 0);
    SplitArgs a;
    a.src = src->ptr; a.stats = src->gn_sums; a.gamma = src->gn_weight; a.beta = src->gn_bias; a.ch = src->channels;
    a.mode = src->mode; a.groups = src->gn_groups; a.count = src->gn_count; a.P = (size_t)H * W; a.planes = (__half*)planes;
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    const size_t total = a.P * (a.ch >> 2);
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)sms * 16) blocks = (size_t)sms * 16;
    split_planes_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}

// ------------------------------------------------------------------------------------------------ state snapshot (training)
// After plane d of the training forward: copy the recurrent state of the five cells -- h_i (the fp32 copies the cells write,
// ping-pong half d & 1) and c_i (tile-blocked, private to the cell kernel) -- into record d + 1 of the per-sample state
// history [D + 1][record] as plain NHWC fp32 (Engine.state_layout: [h0 | c0 | h1 | c1 | ...]).  One launch instead of ten
// strided copies, and graph-friendly: the record index is read from the device-side plane counter (already incremented by
// aarmvs_next_plane) and the history's base address from device memory (one captured graph serves every sample).
namespace {
struct SnapArgs {
    const float* h[5];
    const float* c[5];
    int H, W;
    float* const* state_pp;
    size_t rec;
    int d_host;
    const int* d_dev;
};
__device__ constexpr int SNAP_CH[5] = {16, 16, 16, 16, 8}, SNAP_SCALE[5] = {1, 2, 4, 2, 1};

__global__ void __launch_bounds__(256) snapshot_state_kernel(const SnapArgs a) {
    const int r = plane_index(a.d_host, a.d_dev);            // record to write = planes completed so far
    float* dst = *a.state_pp + (size_t)r * a.rec;
    size_t off = 0;
    for (int i = 0; i < 5; ++i) {
        const int ch = SNAP_CH[i], h = a.H / SNAP_SCALE[i], w = a.W / SNAP_SCALE[i], q = ch / 4;
        const size_t n = (size_t)h * w * ch;
        const int tiles_x = (w + 7) / 8;
        const float* hs = a.h[i] + (size_t)((r - 1) & 1) * n;
        for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
            dst[off + idx] = hs[idx];
            const size_t p = idx / ch;
            const int cn = (int)(idx - p * ch), y = (int)(p / w), x = (int)(p - (size_t)y * w);
            const size_t tile = (size_t)(y >> 4) * tiles_x + (x >> 3);
            const size_t src = ((tile * 4 + cn / q) * 128 + ((y & 15) * 8 + (x & 7))) * q + cn % q;
            dst[off + n + idx] = a.c[i][src];
        }
        off += 2 * n;
    }
}
}  // namespace

extern "C" int aarmvs_snapshot_state(const float* const* h_f32_pairs, const float* const* c_blocked, int H, int W, float* const* state_pp,
                                     size_t record_floats, int d_host, const int* d_dev, void* stream) {
    AARMVS_CHECK_ARG(h_f32_pairs && c_blocked && state_pp && H > 0 && W > 0 && H % 4 == 0 && W % 4 == 0 && record_floats > 0);
    SnapArgs a = {};
    for (int i = 0; i < 5; ++i) {
        AARMVS_CHECK_ARG(h_f32_pairs[i] && c_blocked[i]);
        a.h[i] = h_f32_pairs[i];
        a.c[i] = c_blocked[i];
    }
    a.H = H; a.W = W; a.state_pp = state_pp; a.rec = record_floats; a.d_host = d_host; a.d_dev = d_dev;
    const size_t n = (size_t)H * W * 16;
    const size_t blocks = (n + 255) / 256;
    snapshot_state_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, (cudaStream_t)stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}
