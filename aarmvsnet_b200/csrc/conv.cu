// Direct fp32 convolutions on NHWC tensors: the ConvLSTM cells of the U-Net regulariser (SURVEY.md §8 rows a8,
// a10, a11, a15) and the conv-GN-ReLU trunk of FeatNet (a6, a12).
//
// One templated kernel: a persistent CTA keeps the layer's weights [tap][cin][cout] in shared memory, stages an
// input tile (+halo) gathered from up to three source tensors (channel concat without a `cat`: module.py:80,
// drmvsnet.py:157,161), optionally applying GroupNorm+ReLU of the producer layer or a 2x2 max-pool
// (drmvsnet.py:148,152) while loading, and finishes with a fused epilogue:
//   EPI_RAW    bias, store pre-norm activations, accumulate the GroupNorm sums of this layer in fp64
//   EPI_LSTM   gates i,f,o,g -> c' = sig(f) c + sig(i) tanh(g), h' = sig(o) tanh(c')  (module.py:83-90); c updated in
//              place, h written to the other half of a ping-pong pair (neighbouring tiles still read h_{d-1})
// Thread layout: lane = tile column (conflict-free LDS.128 with a padded channel pitch), a warp owns TCO output
// channels (weights are warp-broadcast LDS.128) and TPX tile rows per lane: TPX*TCO accumulators per thread.
#include "common.cuh"
#include "../../include/aarmvs.h"

enum { EPI_RAW = 0, EPI_LSTM = 1 };

struct ConvSrc {
    const float* ptr;
    const double* stats;
    const float* gamma;
    const float* beta;
    int ch, mode, groups, ch_real;
    double count;
    size_t pp_stride;
    int pp_phase;
};

struct ConvArgs {
    ConvSrc src[3];
    int nsrc;
    int Hin, Win;            // conv input size (after pooling for MAXPOOL2 sources)
    int Hout, Wout;
    const float* weight;     // [KS*KS][CIN][COUT]
    const float* bias;       // [COUT]
    float* out;              // EPI_RAW: [Hout][Wout][cout_real]
    int cout_real;           // EPI_RAW: channels actually stored (<= COUT)
    double* out_stats;       // EPI_RAW: [COUT/8][2] or null
    float* h_pair;           // EPI_LSTM: two [P][CH] halves; plane d writes half d & 1
    float* c_state;          // EPI_LSTM: [P][CH]
    int d_host;
    const int* d_dev;
};

__device__ __forceinline__ const float* src_base(const ConvSrc& S, int d) {
    return S.pp_stride ? S.ptr + (size_t)((d + S.pp_phase) & 1) * S.pp_stride : S.ptr;
}

static int convert_src(const aarmvs_conv_src* in, int nsrc, ConvSrc* out) {
    for (int s = 0; s < nsrc; ++s) {
        if (!in[s].ptr || in[s].channels % 4 != 0 || in[s].channels <= 0) return AARMVS_ERR_ARG;
        out[s].ptr = in[s].ptr;
        out[s].stats = in[s].gn_sums;
        out[s].gamma = in[s].gn_weight;
        out[s].beta = in[s].gn_bias;
        out[s].ch = in[s].channels;
        out[s].mode = in[s].mode;
        out[s].groups = in[s].gn_groups;
        out[s].ch_real = in[s].planes;
        out[s].count = in[s].gn_count;
        out[s].pp_stride = in[s].pp_stride;
        out[s].pp_phase = in[s].pp_phase;
        if (in[s].mode == AARMVS_SRC_GN_RELU &&
            (!in[s].gn_sums || !in[s].gn_weight || !in[s].gn_bias || in[s].gn_groups < 1 || in[s].gn_count <= 0))
            return AARMVS_ERR_ARG;
        if (in[s].mode < 0 || in[s].mode > 3) return AARMVS_ERR_ARG;
    }
    return AARMVS_OK;
}

// GroupNorm scale / shift of input channel c of a GN_RELU source:  y = x * scale + shift
__device__ __forceinline__ void gn_scale_shift(const ConvSrc& S, int cl, float& sc, float& sh) {
    const int gch = S.ch / S.groups;
    float mean, rstd;
    gn_mean_rstd(S.stats + (cl / gch) * 2, S.count, mean, rstd);
    sc = rstd * S.gamma[cl];
    sh = S.beta[cl] - mean * sc;
}

template <int CIN, int COUT, int KS, int STRIDE, int TCO, int TPX, int EPI>
struct ConvCfg {
    static constexpr int NWARPS = 8;
    static constexpr int NCG = COUT / TCO;                 // warps along output channels
    static constexpr int PSETS = NWARPS / NCG;             // warps along tile rows
    static constexpr int TH = PSETS * TPX, TW = 32;
    static constexpr int IH = (TH - 1) * STRIDE + KS, IW = (TW - 1) * STRIDE + KS;
    static constexpr int CPAD = ((CIN / 4) & 1) ? CIN : CIN + 4;   // odd number of 16-byte chunks: conflict-free LDS.128
    static constexpr int W_FLOATS = KS * KS * CIN * COUT;
    static constexpr int IN_FLOATS = IH * IW * CPAD;
    static constexpr size_t SMEM = (size_t)(W_FLOATS + IN_FLOATS + 128) * sizeof(float);
    static_assert(COUT % TCO == 0 && NWARPS % NCG == 0 && TCO % 4 == 0 && CIN % 4 == 0 && CIN <= 64, "bad conv config");
    static_assert(((CPAD / 4) & 1) == 1, "channel pitch must be an odd number of 16-byte chunks");
    static_assert(TCO == 8 || TCO == 16, "a warp owns one or two GroupNorm groups");
};

template <int CIN, int COUT, int KS, int STRIDE, int TCO, int TPX, int EPI>
__global__ void __launch_bounds__(256, 1) conv_kernel(const __grid_constant__ ConvArgs a) {
    using C = ConvCfg<CIN, COUT, KS, STRIDE, TCO, TPX, EPI>;
    extern __shared__ __align__(16) float smem[];
    float* sw = smem;                         // weights
    float* sin = smem + C::W_FLOATS;          // input tile [IH][IW][CPAD]
    float* s_scale = sin + C::IN_FLOATS;      // per input channel GN scale / shift (GN_RELU sources)
    float* s_shift = s_scale + 64;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cg = warp % C::NCG, ps = warp / C::NCG;
    const int co0 = cg * TCO;
    const int d = plane_index(a.d_host, a.d_dev);

    for (int i = tid * 4; i < C::W_FLOATS; i += 256 * 4) st4(sw + i, ldg4(a.weight + i));
    if (tid < CIN) {
        float sc = 1.f, sh = 0.f;
        int base = 0;
        for (int s = 0; s < a.nsrc; ++s) {
            if (tid >= base && tid < base + a.src[s].ch && a.src[s].mode == AARMVS_SRC_GN_RELU)
                gn_scale_shift(a.src[s], tid - base, sc, sh);
            base += a.src[s].ch;
        }
        s_scale[tid] = sc;
        s_shift[tid] = sh;
    }
    const float* sptr[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) sptr[s] = s < a.nsrc ? src_base(a.src[s], d) : nullptr;
    __syncthreads();

    const int tiles_x = (a.Wout + C::TW - 1) / C::TW, tiles_y = (a.Hout + C::TH - 1) / C::TH;
    float gsum[TCO / 8] = {}, gsq[TCO / 8] = {};          // per-thread GroupNorm partial sums (groups of 8 channels)

    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int oy0 = ty * C::TH, ox0 = tx * C::TW;
        const int iy0 = oy0 * STRIDE - (KS - 1) / 2, ix0 = ox0 * STRIDE - (KS - 1) / 2;
        // ---- stage the input tile: item = (position, 16-byte channel chunk)
        constexpr int CIN4 = CIN / 4;
        for (int it = tid; it < C::IH * C::IW * CIN4; it += 256) {
            const int pos = it / CIN4, c4 = it - pos * CIN4;
            const int r = pos / C::IW, cc = pos - r * C::IW;
            const int y = iy0 + r, x = ix0 + cc;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < a.Hin && x >= 0 && x < a.Win) {
                const int c = c4 * 4;
                int base = 0, s = 0;
                while (s < a.nsrc - 1 && c >= base + a.src[s].ch) { base += a.src[s].ch; ++s; }
                const ConvSrc& S = a.src[s];
                const float* sp = s == 0 ? sptr[0] : s == 1 ? sptr[1] : sptr[2];
                const int cl = c - base;
                if (EPI == EPI_LSTM && d == 0 && s == a.nsrc - 1) {
                    // h_{-1} = 0 (module.py:94-96): the state buffers are never read on the first plane
                } else if (S.mode == AARMVS_SRC_MAXPOOL2) {
                    const size_t rowp = (size_t)2 * a.Win * S.ch;
                    const float* p = sp + (size_t)(2 * y) * rowp + (size_t)(2 * x) * S.ch + cl;
                    float4 q0 = ld4(p), q1 = ld4(p + S.ch), q2 = ld4(p + rowp), q3 = ld4(p + rowp + S.ch);
                    v.x = fmaxf(fmaxf(q0.x, q1.x), fmaxf(q2.x, q3.x));
                    v.y = fmaxf(fmaxf(q0.y, q1.y), fmaxf(q2.y, q3.y));
                    v.z = fmaxf(fmaxf(q0.z, q1.z), fmaxf(q2.z, q3.z));
                    v.w = fmaxf(fmaxf(q0.w, q1.w), fmaxf(q2.w, q3.w));
                } else if (S.mode == AARMVS_SRC_NCHW) {
                    const size_t plane = (size_t)a.Hin * a.Win, o = (size_t)y * a.Win + x;
                    v.x = cl + 0 < S.ch_real ? __ldg(sp + (cl + 0) * plane + o) : 0.f;
                    v.y = cl + 1 < S.ch_real ? __ldg(sp + (cl + 1) * plane + o) : 0.f;
                    v.z = cl + 2 < S.ch_real ? __ldg(sp + (cl + 2) * plane + o) : 0.f;
                    v.w = cl + 3 < S.ch_real ? __ldg(sp + (cl + 3) * plane + o) : 0.f;
                } else {
                    v = ld4(sp + ((size_t)y * a.Win + x) * S.ch + cl);
                    if (S.mode == AARMVS_SRC_GN_RELU) {
                        v.x = fmaxf(fmaf(v.x, s_scale[c + 0], s_shift[c + 0]), 0.f);
                        v.y = fmaxf(fmaf(v.y, s_scale[c + 1], s_shift[c + 1]), 0.f);
                        v.z = fmaxf(fmaf(v.z, s_scale[c + 2], s_shift[c + 2]), 0.f);
                        v.w = fmaxf(fmaf(v.w, s_scale[c + 3], s_shift[c + 3]), 0.f);
                    }
                }
            }
            st4(sin + pos * C::CPAD + c4 * 4, v);
        }
        __syncthreads();

        // ---- main loop
        float acc[TPX][TCO];
#pragma unroll
        for (int p = 0; p < TPX; ++p)
#pragma unroll
            for (int k = 0; k < TCO; ++k) acc[p][k] = 0.f;
#pragma unroll 1
        for (int tap = 0; tap < KS * KS; ++tap) {
            const int ky = tap / KS, kx = tap - ky * KS;
            const float* wt = sw + tap * CIN * COUT + co0;
            const float* it = sin + ((ps * TPX * STRIDE + ky) * C::IW + lane * STRIDE + kx) * C::CPAD;
#pragma unroll 2
            for (int c4 = 0; c4 < CIN4; ++c4) {
                float4 x[TPX];
#pragma unroll
                for (int p = 0; p < TPX; ++p) x[p] = ld4(it + p * STRIDE * C::IW * C::CPAD + c4 * 4);
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    float wv[TCO];
#pragma unroll
                    for (int k = 0; k < TCO; k += 4) {
                        float4 w4 = ld4(wt + (c4 * 4 + ci) * COUT + k);
                        wv[k] = w4.x; wv[k + 1] = w4.y; wv[k + 2] = w4.z; wv[k + 3] = w4.w;
                    }
#pragma unroll
                    for (int p = 0; p < TPX; ++p) {
                        const float xv = ci == 0 ? x[p].x : ci == 1 ? x[p].y : ci == 2 ? x[p].z : x[p].w;
#pragma unroll
                        for (int k = 0; k < TCO; ++k) acc[p][k] = fmaf(xv, wv[k], acc[p][k]);
                    }
                }
            }
        }

        // ---- epilogue
        const int ox = ox0 + lane;
#pragma unroll
        for (int p = 0; p < TPX; ++p) {
            const int oy = oy0 + ps * TPX + p;
            if (oy >= a.Hout || ox >= a.Wout) continue;
            const size_t pix = (size_t)oy * a.Wout + ox;
            if (EPI == EPI_RAW) {
#pragma unroll
                for (int k = 0; k < TCO; k += 4) {
                    float4 o = make_float4(acc[p][k] + a.bias[co0 + k], acc[p][k + 1] + a.bias[co0 + k + 1],
                                           acc[p][k + 2] + a.bias[co0 + k + 2], acc[p][k + 3] + a.bias[co0 + k + 3]);
                    if (co0 + k + 3 < a.cout_real) st4(a.out + pix * a.cout_real + co0 + k, o);
                    gsum[k / 8] += (o.x + o.y) + (o.z + o.w);
                    gsq[k / 8] += (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
                }
            } else {   // EPI_LSTM: this thread holds the gates of NH hidden channels, packed [hidden channel][i,f,o,g]
                constexpr int CH = COUT / 4, NH = TCO / 4;
                float* hw = a.h_pair + (size_t)(d & 1) * a.Hout * a.Wout * CH + pix * CH + cg * NH;
                float* cp = a.c_state + pix * CH + cg * NH;
                float cprev[NH], hn[NH], cn[NH];
#pragma unroll
                for (int j = 0; j < NH; j += 4) {
                    float4 t = d == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4(cp + j);
                    cprev[j] = t.x; cprev[j + 1] = t.y; cprev[j + 2] = t.z; cprev[j + 3] = t.w;
                }
#pragma unroll
                for (int j = 0; j < NH; ++j) {
                    const float gi = sigmoidf_acc(acc[p][j * 4 + 0] + a.bias[co0 + j * 4 + 0]);
                    const float gf = sigmoidf_acc(acc[p][j * 4 + 1] + a.bias[co0 + j * 4 + 1]);
                    const float go = sigmoidf_acc(acc[p][j * 4 + 2] + a.bias[co0 + j * 4 + 2]);
                    const float gg = tanhf(acc[p][j * 4 + 3] + a.bias[co0 + j * 4 + 3]);
                    cn[j] = __fadd_rn(__fmul_rn(gf, cprev[j]), __fmul_rn(gi, gg));
                    hn[j] = go * tanhf(cn[j]);
                }
#pragma unroll
                for (int j = 0; j < NH; j += 4) {
                    st4(cp + j, make_float4(cn[j], cn[j + 1], cn[j + 2], cn[j + 3]));
                    st4(hw + j, make_float4(hn[j], hn[j + 1], hn[j + 2], hn[j + 3]));
                }
            }
        }
        __syncthreads();      // the tile is reused by the next iteration
    }

    if (EPI == EPI_RAW && a.out_stats) {
#pragma unroll
        for (int g = 0; g < TCO / 8; ++g) {
            double s = warp_sum((double)gsum[g]), q = warp_sum((double)gsq[g]);
            if (lane == 0) {
                atomicAdd(a.out_stats + (co0 / 8 + g) * 2 + 0, s);
                atomicAdd(a.out_stats + (co0 / 8 + g) * 2 + 1, q);
            }
        }
    }
}

// multiprocessor count of the CURRENT device (a cached device attribute of the runtime: no global state here, and a process
// that drives several GPUs gets each one's own count)
static int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return -1;
    return n;
}

extern "C" int aarmvs_version(void) { return 5; }
extern "C" int aarmvs_sm_count(void) { return sm_count(); }

template <int CIN, int COUT, int KS, int STRIDE, int TCO, int TPX, int EPI>
static int launch_conv(const ConvArgs& a, cudaStream_t stream) {
    using C = ConvCfg<CIN, COUT, KS, STRIDE, TCO, TPX, EPI>;
    auto kern = conv_kernel<CIN, COUT, KS, STRIDE, TCO, TPX, EPI>;
    static unsigned long long smem_ok = 0;          // per (kernel, device): the attribute is a per-device property
    if (cudaError_t e = ensure_dyn_smem(kern, (int)C::SMEM, smem_ok)) return (int)e;
    const int tiles = ((a.Wout + C::TW - 1) / C::TW) * ((a.Hout + C::TH - 1) / C::TH);
    const int sms = sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    kern<<<tiles < sms ? tiles : sms, 256, C::SMEM, stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}

// ---- FeatNet / generic conv + GroupNorm sums ----------------------------------------------------------------
extern "C" int aarmvs_conv2d(const aarmvs_conv_src* src, int nsrc, int Hin, int Win, int cin, int cout, int ksize,
                             int stride, const float* weight, const float* bias, float* out, int cout_stored,
                             double* out_gn_sums, void* stream) {
    AARMVS_CHECK_ARG(src && weight && bias && out && Hin > 0 && Win > 0 && nsrc >= 1 && nsrc <= 3);
    AARMVS_CHECK_ARG(cout_stored > 0 && cout_stored <= cout && cout_stored % 4 == 0);
    ConvArgs a = {};
    a.nsrc = nsrc;
    if (int e = convert_src(src, nsrc, a.src)) return e;
    int csum = 0;
    for (int s = 0; s < nsrc; ++s) csum += src[s].channels;
    AARMVS_CHECK_ARG(csum == cin);
    const int pad = (ksize - 1) / 2;
    a.Hin = Hin; a.Win = Win;
    a.Hout = (Hin + 2 * pad - ksize) / stride + 1;
    a.Wout = (Win + 2 * pad - ksize) / stride + 1;
    a.weight = weight; a.bias = bias; a.out = out; a.cout_real = cout_stored; a.out_stats = out_gn_sums;
    cudaStream_t st = (cudaStream_t)stream;
#define CONV_CASE(CI, CO, K, S, TCO, TPX) \
    if (cin == CI && cout == CO && ksize == K && stride == S) return launch_conv<CI, CO, K, S, TCO, TPX, EPI_RAW>(a, st);
    CONV_CASE(4, 8, 3, 1, 8, 2)        // init_conv.0 (RGB padded to 4 planes)
    CONV_CASE(8, 16, 3, 1, 16, 2)      // init_conv.1
    CONV_CASE(16, 32, 3, 1, 16, 4)     // conv0
    CONV_CASE(32, 32, 3, 2, 16, 2)     // conv1, conv2 (stride 2)
    CONV_CASE(32, 32, 3, 1, 16, 4)     // p_conv + m_conv fused (27 -> padded 32 output channels)
    CONV_CASE(32, 16, 1, 1, 16, 2)     // intraAA.conv0 (1x1)
    CONV_CASE(32, 8, 1, 1, 8, 2)       // intraAA.conv1 / conv2 (1x1)
#undef CONV_CASE
    return AARMVS_ERR_ARG;
}

// ---- ConvLSTM cell ------------------------------------------------------------------------------------------
extern "C" int aarmvs_convlstm_cell(const aarmvs_conv_src* src, int nsrc, int H, int W, int cin, int hidden,
                                    const float* weight_packed, const float* bias_packed, float* h_pair, float* c_state,
                                    int d_host, const int* d_dev, void* stream) {
    AARMVS_CHECK_ARG(src && weight_packed && bias_packed && h_pair && c_state && H > 0 && W > 0);
    AARMVS_CHECK_ARG(nsrc >= 1 && nsrc <= 2);
    aarmvs_conv_src all[3];
    for (int s = 0; s < nsrc; ++s) all[s] = src[s];
    all[nsrc] = aarmvs_conv_src{};
    all[nsrc].ptr = h_pair;                       // h_{d-1}: the half plane d does not write
    all[nsrc].channels = hidden;
    all[nsrc].mode = AARMVS_SRC_PLAIN;
    all[nsrc].pp_stride = (size_t)H * W * hidden;
    all[nsrc].pp_phase = 1;
    ConvArgs a = {};
    a.nsrc = nsrc + 1;
    if (int e = convert_src(all, nsrc + 1, a.src)) return e;
    int csum = 0;
    for (int s = 0; s <= nsrc; ++s) csum += all[s].channels;
    AARMVS_CHECK_ARG(csum == cin);
    a.Hin = a.Hout = H; a.Win = a.Wout = W;
    a.weight = weight_packed; a.bias = bias_packed; a.h_pair = h_pair; a.c_state = c_state;
    a.d_host = d_host; a.d_dev = d_dev;
    cudaStream_t st = (cudaStream_t)stream;
    if (cin == 48 && hidden == 16) return launch_conv<48, 64, 3, 1, 16, 4, EPI_LSTM>(a, st);   // cells 0, 3
    if (cin == 32 && hidden == 16) return launch_conv<32, 64, 3, 1, 16, 4, EPI_LSTM>(a, st);   // cells 1, 2
    if (cin == 40 && hidden == 8) return launch_conv<40, 32, 3, 1, 16, 4, EPI_LSTM>(a, st);    // cell 4
    return AARMVS_ERR_ARG;
}

// ---- deConvGnReLU's transposed conv (module.py:281): 16 -> 16 channels, k3 s2 p1 op1 --------------------------
// out[2i+a][2j+b] gathers the input pixels (i,j), (i,j+1), (i+1,j), (i+1,j+1) with the taps
//   (0,0): in(i,j) W[1][1]
//   (0,1): in(i,j) W[1][2] + in(i,j+1) W[1][0]
//   (1,0): in(i,j) W[2][1] + in(i+1,j) W[0][1]
//   (1,1): in(i,j) W[2][2] + in(i,j+1) W[2][0] + in(i+1,j) W[0][2] + in(i+1,j+1) W[0][0]
// (from y = 2 i - 1 + ky).  One thread = one input pixel = one 2x2 output quad, all 16 output channels.
struct DeconvArgs {
    ConvSrc in;
    int Hin, Win;
    float* out;
    double* out_stats;     // 2 groups
    int d_host;
    const int* d_dev;
};

struct DeconvW {           // by value in the kernel parameter (constant) bank: every FMA takes its weight operand from
    float w[9 * 256];      // c[0][...] with no load instruction ([ky][kx][cin][cout])
    float b[16];
};

// acc[co] += sum_ci x[ci] * w[ci][co]   (w: compile-time offset into the parameter bank)
__device__ __forceinline__ void deconv_acc(float (&acc)[16], const float (&x)[16], const float* w) {
#pragma unroll
    for (int ci = 0; ci < 16; ++ci)
#pragma unroll
        for (int k = 0; k < 16; ++k) acc[k] = fmaf(x[ci], w[ci * 16 + k], acc[k]);
}

__global__ void __launch_bounds__(128) deconv2x_kernel(const __grid_constant__ DeconvArgs a, const __grid_constant__ DeconvW W) {
    __shared__ float s_scale[16], s_shift[16];
    if (threadIdx.x < 16) {
        float sc = 1.f, sh = 0.f;
        if (a.in.mode == AARMVS_SRC_GN_RELU) gn_scale_shift(a.in, threadIdx.x, sc, sh);
        s_scale[threadIdx.x] = sc;
        s_shift[threadIdx.x] = sh;
    }
    __syncthreads();
    const int d = plane_index(a.d_host, a.d_dev);
    const float* in = src_base(a.in, d);
    const int lane = threadIdx.x & 31;
    const int n = a.Hin * a.Win, Wo = 2 * a.Win;
    float gs[2] = {0.f, 0.f}, gq[2] = {0.f, 0.f};
    for (int p0 = blockIdx.x * blockDim.x; p0 < n; p0 += gridDim.x * blockDim.x) {
        const int p = p0 + threadIdx.x;
        if (p < n) {
            const int i = p / a.Win, j = p - i * a.Win;
            float x00[16], x01[16], x10[16], x11[16];
            auto load = [&](float (&x)[16], int yy, int xx) {
                const bool ok = yy < a.Hin && xx < a.Win;
#pragma unroll
                for (int k = 0; k < 16; k += 4) {
                    float4 v = ok ? ldg4(in + ((size_t)yy * a.Win + xx) * 16 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok && a.in.mode == AARMVS_SRC_GN_RELU) {
                        v.x = fmaxf(fmaf(v.x, s_scale[k], s_shift[k]), 0.f);
                        v.y = fmaxf(fmaf(v.y, s_scale[k + 1], s_shift[k + 1]), 0.f);
                        v.z = fmaxf(fmaf(v.z, s_scale[k + 2], s_shift[k + 2]), 0.f);
                        v.w = fmaxf(fmaf(v.w, s_scale[k + 3], s_shift[k + 3]), 0.f);
                    }
                    x[k] = v.x; x[k + 1] = v.y; x[k + 2] = v.z; x[k + 3] = v.w;
                }
            };
            load(x00, i, j); load(x01, i, j + 1); load(x10, i + 1, j); load(x11, i + 1, j + 1);
#pragma unroll
            for (int row = 0; row < 2; ++row) {        // output row 2i + row: the two columns 2j, 2j+1 are 128 contiguous bytes
                float acc[2][16];
#pragma unroll
                for (int k = 0; k < 16; ++k) acc[0][k] = acc[1][k] = W.b[k];
                if (row == 0) {
                    deconv_acc(acc[0], x00, W.w + (1 * 3 + 1) * 256);
                    deconv_acc(acc[1], x00, W.w + (1 * 3 + 2) * 256);
                    deconv_acc(acc[1], x01, W.w + (1 * 3 + 0) * 256);
                } else {
                    deconv_acc(acc[0], x00, W.w + (2 * 3 + 1) * 256);
                    deconv_acc(acc[0], x10, W.w + (0 * 3 + 1) * 256);
                    deconv_acc(acc[1], x00, W.w + (2 * 3 + 2) * 256);
                    deconv_acc(acc[1], x01, W.w + (2 * 3 + 0) * 256);
                    deconv_acc(acc[1], x10, W.w + (0 * 3 + 2) * 256);
                    deconv_acc(acc[1], x11, W.w + (0 * 3 + 0) * 256);
                }
                float* o = a.out + ((size_t)(2 * i + row) * Wo + 2 * j) * 16;
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k = 0; k < 16; k += 4) {
                        st4(o + c * 16 + k, make_float4(acc[c][k], acc[c][k + 1], acc[c][k + 2], acc[c][k + 3]));
                        gs[k / 8] += (acc[c][k] + acc[c][k + 1]) + (acc[c][k + 2] + acc[c][k + 3]);
                        gq[k / 8] += (acc[c][k] * acc[c][k] + acc[c][k + 1] * acc[c][k + 1]) + (acc[c][k + 2] * acc[c][k + 2] + acc[c][k + 3] * acc[c][k + 3]);
                    }
            }
        }
    }
    if (a.out_stats) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            double s = warp_sum((double)gs[g]), q = warp_sum((double)gq[g]);
            if (lane == 0) {
                atomicAdd(a.out_stats + g * 2 + 0, s);
                atomicAdd(a.out_stats + g * 2 + 1, q);
            }
        }
    }
}

extern "C" int aarmvs_deconv2x(const aarmvs_conv_src* in, int Hin, int Win, const float* weight_host, const float* bias_host,
                               float* out, double* out_gn_sums, int d_host, const int* d_dev, void* stream) {
    AARMVS_CHECK_ARG(in && weight_host && bias_host && out && Hin > 0 && Win > 0 && in->channels == 16);
    AARMVS_CHECK_ARG(in->mode == AARMVS_SRC_PLAIN || in->mode == AARMVS_SRC_GN_RELU);
    DeconvArgs a = {};
    if (int e = convert_src(in, 1, &a.in)) return e;
    a.Hin = Hin; a.Win = Win; a.out = out; a.out_stats = out_gn_sums;
    a.d_host = d_host; a.d_dev = d_dev;
    DeconvW w;
    for (int i = 0; i < 9 * 256; ++i) w.w[i] = weight_host[i];
    for (int i = 0; i < 16; ++i) w.b[i] = bias_host[i];
    const int n = Hin * Win, sms = sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    int blocks = (n + 127) / 128;
    if (blocks > sms * 8) blocks = sms * 8;
    deconv2x_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(a, w);
    AARMVS_LAUNCH_RESULT();
}

// ---- cost head (conv_0: 3x3, 8 -> 1) + winner-take-all update (drmvsnet.py:165, 324-339) ----------------------
struct HeadArgs {
    ConvSrc h4;
    int H, W;
    float w[72];           // [tap][8]
    float bias;
    const float* depth_values;
    int num_depth, d_host;
    const int* d_dev;
    float *max_prob, *depth_map, *exp_sum, *conf, *cost_volume;
    int* index_map;
};

__global__ void __launch_bounds__(256) head_wta_kernel(const __grid_constant__ HeadArgs a) {
    const int d = plane_index(a.d_host, a.d_dev);
    const float* h = src_base(a.h4, d);
    const float depth = a.depth_values[d];
    const int n = a.H * a.W;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const int y = p / a.W, x = p - y * a.W;
        float acc = a.bias;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int yy = y + ky - 1;
            if (yy < 0 || yy >= a.H) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int xx = x + kx - 1;
                if (xx < 0 || xx >= a.W) continue;
                const float* q = h + ((size_t)yy * a.W + xx) * 8;
                const float4 v0 = ld4(q), v1 = ld4(q + 4);
                const float* w = a.w + (ky * 3 + kx) * 8;
                acc = fmaf(v0.x, w[0], acc); acc = fmaf(v0.y, w[1], acc); acc = fmaf(v0.z, w[2], acc); acc = fmaf(v0.w, w[3], acc);
                acc = fmaf(v1.x, w[4], acc); acc = fmaf(v1.y, w[5], acc); acc = fmaf(v1.z, w[6], acc); acc = fmaf(v1.w, w[7], acc);
            }
        }
        if (a.cost_volume) a.cost_volume[(size_t)d * n + p] = acc;
        const float prob = expf(acc);
        float mp = d == 0 ? 0.f : a.max_prob[p];
        float es = d == 0 ? 0.f : a.exp_sum[p];
        if (d == 0) { a.depth_map[p] = 0.f; a.index_map[p] = 0; }
        if (mp < prob) {           // strict '<': the first plane wins ties (drmvsnet.py:327)
            mp = prob;
            a.depth_map[p] = depth;
            a.index_map[p] = d;
        }
        es = __fadd_rn(es, prob);
        a.max_prob[p] = mp;
        a.exp_sum[p] = es;
        if (d == a.num_depth - 1) a.conf[p] = __fdiv_rn(mp, es);
    }
}

extern "C" int aarmvs_head_wta(const aarmvs_conv_src* h4, int H, int W, const float* weight_host, const float* bias_host,
                               const float* depth_values, int num_depth, int d_host, const int* d_dev, float* max_prob,
                               float* depth_map, float* exp_sum, int* index_map, float* conf, float* cost_volume,
                               void* stream) {
    AARMVS_CHECK_ARG(h4 && weight_host && bias_host && depth_values && max_prob && depth_map && exp_sum && index_map && conf);
    AARMVS_CHECK_ARG(H > 0 && W > 0 && num_depth > 0 && h4->channels == 8 && h4->mode == AARMVS_SRC_PLAIN);
    HeadArgs a = {};
    if (int e = convert_src(h4, 1, &a.h4)) return e;
    a.H = H; a.W = W;
    for (int i = 0; i < 72; ++i) a.w[i] = weight_host[i];
    a.bias = bias_host[0];
    a.depth_values = depth_values; a.num_depth = num_depth; a.d_host = d_host; a.d_dev = d_dev;
    a.max_prob = max_prob; a.depth_map = depth_map; a.exp_sum = exp_sum; a.index_map = index_map; a.conf = conf;
    a.cost_volume = cost_volume;
    const int n = H * W, sms = sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    int blocks = (n + 255) / 256;
    if (blocks > sms * 8) blocks = sms * 8;
    head_wta_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}

struct NextPlaneArgs { int* d_dev; double* sums; int nzero; };
__global__ void next_plane_kernel(const NextPlaneArgs a) {
    int* d_dev = a.d_dev; double* sums = a.sums; const int nzero = a.nzero;
    asm volatile("griddepcontrol.wait;" ::: "memory");           // part of the per-plane PDL chain: the head kernel is complete
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int i = threadIdx.x; i < nzero; i += blockDim.x) sums[i] = 0.0;
    if (threadIdx.x == 0 && d_dev) *d_dev += 1;
}

extern "C" int aarmvs_next_plane(int* d_dev, double* sums, int nzero, void* stream) {
    AARMVS_CHECK_ARG(nzero >= 0 && (nzero == 0 || sums));
    NextPlaneArgs a = {d_dev, sums, nzero};
    if (cudaError_t e = launch_pdl(next_plane_kernel, dim3(1), dim3(128), 0, (cudaStream_t)stream, a)) return (int)e;
    AARMVS_LAUNCH_RESULT();
}
