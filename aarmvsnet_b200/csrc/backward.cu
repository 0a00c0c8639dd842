// Backward pass of the path, part 1: the dense primitives (BASELINE config 5; train.py:288-307 `loss.backward()` through
// models/drmvsnet.py:272-295).  Everything here is fp32 on NHWC tensors, CUDA-core FMA (the cells' data and weight gradients
// run on the tensor cores: conv_dx_tc.cu, conv_dw_tc.cu; conv_dx3_kernel / conv_wgrad3_kernel below are their fp32 fallbacks):
//
//   conv_nhwc_kernel      direct convolution / transposed convolution over a channel CONCAT of inputs, writing (or adding
//                         to) a channel SPLIT of outputs.  One kernel covers the data gradients dX = dY (*) W^T of the small
//                         and the strided layers (the deconvolution's is a strided conv, module.py:281), the ConvTranspose
//                         forward itself, and the 1x1 layers of FeatNet.
//   conv_dx3_kernel       the same sum for the five cells' 3x3 data gradients, register tiled (the widest convolutions).
//   conv_wgrad_kernel     dW[tap][ci][co] += sum_p X[p + tap][ci] * dY[p][co] (+ bias gradient), fp32 atomics at the end.
//   conv_wgrad3_kernel    the same for 3x3 stride-1 layers: one staged tile for all nine taps, a warp pair per tap;
//   conv_wgrad3_narrow_   ... and for 1 / 4 output channels (head, omega's stem).
//   gn_*                  GroupNorm (+ReLU) materialisation, statistics, and the two-pass backward (global sums, then dX).
//   lstm_gates_bwd(_act)  module.py:83-90 backwards: d gates, d c_{d-1} from d h_d, d c_d (from the gates the forward stored).
//   maxpool2 / _bwd       drmvsnet.py:148,152; the gradient goes to the first maximum of the window in scan order (ATen).
//
// The training shape is small (128 x 160 feature maps, B = 4, D = 192: SURVEY.md §8d cfg5), so the backward keeps whole
// fp32 activations in HBM and is bound by the ~60 dependent launches of a plane, not by the tensor pipe or HBM.
#include <type_traits>

#include "common.cuh"
#include "../../include/aarmvs.h"

namespace {

struct Seg {
    float* ptr;
    int ch;
    int acc;
};

static int convert_segs(const aarmvs_seg* in, int n, Seg* out, int& total) {
    total = 0;
    for (int s = 0; s < n; ++s) {
        if (!in[s].ptr || in[s].channels <= 0) return AARMVS_ERR_ARG;
        out[s].ptr = in[s].ptr;
        out[s].ch = in[s].channels;
        out[s].acc = in[s].accumulate;
        total += in[s].channels;
    }
    return AARMVS_OK;
}

// ------------------------------------------------------------------------------------------------ generic conv
constexpr int GC_THREADS = 128;     // one thread = one output pixel x GC_CO output channels (8 for small maps: twice the CTAs)

struct GConvArgs {
    Seg in[3], out[3];
    int nin, nout;
    int Hin, Win, Hout, Wout, cin, cout, ks, stride, transposed;
    const float* w;        // [ks*ks][cin][cout]
    const float* bias;     // [cout] or null
};

template <int GC_CO>
__global__ void __launch_bounds__(GC_THREADS) conv_nhwc_kernel(const __grid_constant__ GConvArgs a) {
    extern __shared__ __align__(16) float sw[];          // [ks*ks][cin][GC_CO] of this output-channel tile, zero padded
    const int tid = threadIdx.x;
    const int co0 = blockIdx.y * GC_CO;
    const int rows = a.ks * a.ks * a.cin;
    for (int i = tid; i < rows * GC_CO; i += GC_THREADS) {
        const int r = i / GC_CO, k = i - r * GC_CO;
        sw[i] = co0 + k < a.cout ? __ldg(a.w + (size_t)r * a.cout + co0 + k) : 0.f;
    }
    __syncthreads();
    const int P = a.Hout * a.Wout, pad = (a.ks - 1) / 2;
    for (int p = blockIdx.x * GC_THREADS + tid; p < P; p += gridDim.x * GC_THREADS) {
        const int Y = p / a.Wout, X = p - Y * a.Wout;
        float acc[GC_CO];
#pragma unroll
        for (int k = 0; k < GC_CO; ++k) acc[k] = (a.bias && co0 + k < a.cout) ? __ldg(a.bias + co0 + k) : 0.f;
        for (int ky = 0; ky < a.ks; ++ky) {
            int y;
            if (a.transposed) {
                const int t = Y + pad - ky;
                if (t < 0 || t % a.stride) continue;
                y = t / a.stride;
            } else {
                y = a.stride * Y - pad + ky;
            }
            if (y < 0 || y >= a.Hin) continue;
            for (int kx = 0; kx < a.ks; ++kx) {
                int x;
                if (a.transposed) {
                    const int t = X + pad - kx;
                    if (t < 0 || t % a.stride) continue;
                    x = t / a.stride;
                } else {
                    x = a.stride * X - pad + kx;
                }
                if (x < 0 || x >= a.Win) continue;
                const float* wt = sw + (size_t)(ky * a.ks + kx) * a.cin * GC_CO;
                const size_t pix = (size_t)y * a.Win + x;
                for (int s = 0; s < a.nin; ++s) {
                    const int ch = a.in[s].ch;
                    const float* ip = a.in[s].ptr + pix * ch;
                    if ((ch & 3) == 0) {
                        for (int c = 0; c < ch; c += 4) {
                            const float4 v = ld4(ip + c);
                            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float* wr = wt + (c + e) * GC_CO;
#pragma unroll
                                for (int k = 0; k < GC_CO; k += 4) {
                                    const float4 w4 = ld4(wr + k);
                                    acc[k] = fmaf(vv[e], w4.x, acc[k]);
                                    acc[k + 1] = fmaf(vv[e], w4.y, acc[k + 1]);
                                    acc[k + 2] = fmaf(vv[e], w4.z, acc[k + 2]);
                                    acc[k + 3] = fmaf(vv[e], w4.w, acc[k + 3]);
                                }
                            }
                        }
                    } else {
                        for (int c = 0; c < ch; ++c) {
                            const float v = ip[c];
                            const float* wr = wt + c * GC_CO;
#pragma unroll
                            for (int k = 0; k < GC_CO; ++k) acc[k] = fmaf(v, wr[k], acc[k]);
                        }
                    }
                    wt += ch * GC_CO;
                }
            }
        }
        // store: output channel co0 + k lives in the segment that covers it
        int base = 0;
        for (int s = 0; s < a.nout; ++s) {
            const int ch = a.out[s].ch;
            float* op = a.out[s].ptr + (size_t)p * ch;
#pragma unroll
            for (int k = 0; k < GC_CO; ++k) {
                const int co = co0 + k;
                if (co >= base && co < base + ch && co < a.cout) {
                    float* q = op + (co - base);
                    *q = a.out[s].acc ? *q + acc[k] : acc[k];
                }
            }
            base += ch;
        }
    }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// Block = 256 threads = 4 pixel groups x (4 input-channel quads x 16 output-channel quads): every thread owns a 4 x 4 block of
// dW[tap][ci0 .. ci0 + 16][0 .. 64) for a quarter of the pixels of each staged tile; tiles of 64 pixels are staged by all 256
// threads (16-byte loads where the channel counts allow), partial sums go to global memory with fp32 atomics at the end.
// (The first version -- 64 threads, 512 pixels per block, scalar staging -- took ~100 us per call whatever the layer: latency.)
constexpr int WG_PX = 64, WG_CI = 16, WG_CO = 64, WG_THREADS = 256;

struct WGArgs {
    Seg x[3];
    int nx;
    const float* dy;       // [Hy][Wy][cy]
    int Hx, Wx, Hy, Wy, cx, cy, ks, stride, chunk;
    float* dw;             // [ks*ks][cx][cy], +=
    float* db;             // [cy], += (or null)
};

__global__ void __launch_bounds__(WG_THREADS) conv_wgrad_kernel(const __grid_constant__ WGArgs a) {
    __shared__ __align__(16) float xs[WG_PX][WG_CI];
    __shared__ __align__(16) float ds[WG_PX][WG_CO];
    const int tid = threadIdx.x;
    const int tap = blockIdx.y, ky = tap / a.ks, kx = tap - ky * a.ks, pad = (a.ks - 1) / 2;
    const int ci0 = blockIdx.z * WG_CI;
    const int pg = tid >> 6, ci4 = (tid >> 4) & 3, co4 = tid & 15;
    const int P = a.Hy * a.Wy;
    const int p_begin = blockIdx.x * a.chunk, p_end = min(P, p_begin + a.chunk);
    const int cy4 = (a.cy + 3) & ~3;                                     // output channels staged (zero padded to a quad)
    const bool dy_vec = (a.cy & 3) == 0;
    float acc[4][4] = {};
    float bsum[4] = {0.f, 0.f, 0.f, 0.f};
    const bool active = co4 * 4 < a.cy;
    const bool want_bias = a.db && tap == 0 && blockIdx.z == 0 && ci4 == 0;
    // the segment (and channel inside it) this thread stages: constant over the tiles
    const int sc = ci0 + (tid & 3) * 4;                                  // first of the 4 channels of this thread's staging quad
    int sseg = 0, sbase = 0;
    while (sseg < a.nx - 1 && sc >= sbase + a.x[sseg].ch) { sbase += a.x[sseg].ch; ++sseg; }
    const bool x_vec = (a.x[sseg].ch & 3) == 0 && (sbase & 3) == 0 && sc + 3 < a.cx && (sc - sbase) + 3 < a.x[sseg].ch;
    for (int p0 = p_begin; p0 < p_end; p0 += WG_PX) {
        {   // stage X[p + tap][ci0 .. ci0 + 16): thread = (pixel tid / 4, channel quad tid % 4)
            const int px = tid >> 2, p = p0 + px;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < p_end && sc < a.cx) {
                const int Y = p / a.Wy, X = p - Y * a.Wy;
                const int y = a.stride * Y - pad + ky, x = a.stride * X - pad + kx;
                if (y >= 0 && y < a.Hx && x >= 0 && x < a.Wx) {
                    const size_t pix = (size_t)y * a.Wx + x;
                    if (x_vec) {
                        v = ld4(a.x[sseg].ptr + pix * a.x[sseg].ch + (sc - sbase));
                    } else {
                        float e[4] = {0.f, 0.f, 0.f, 0.f};
                        for (int k = 0; k < 4; ++k) {
                            const int c = sc + k;
                            if (c >= a.cx) break;
                            int base = 0, sg = 0;
                            while (sg < a.nx - 1 && c >= base + a.x[sg].ch) { base += a.x[sg].ch; ++sg; }
                            e[k] = a.x[sg].ptr[pix * a.x[sg].ch + (c - base)];
                        }
                        v = make_float4(e[0], e[1], e[2], e[3]);
                    }
                }
            }
            *reinterpret_cast<float4*>(&xs[px][(tid & 3) * 4]) = v;
        }
        for (int it = tid; it < WG_PX * (cy4 >> 2); it += WG_THREADS) {  // stage dY[p][0 .. cy4)
            const int px = it / (cy4 >> 2), q = it - px * (cy4 >> 2);
            const int p = p0 + px;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p < p_end) {
                const float* dp = a.dy + (size_t)p * a.cy + q * 4;
                if (dy_vec) v = ld4(dp);
                else {
                    float e[4] = {0.f, 0.f, 0.f, 0.f};
                    for (int k = 0; k < 4 && q * 4 + k < a.cy; ++k) e[k] = dp[k];
                    v = make_float4(e[0], e[1], e[2], e[3]);
                }
            }
            *reinterpret_cast<float4*>(&ds[px][q * 4]) = v;
        }
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int j = 0; j < WG_PX / 4; ++j) {
                const int px = pg * (WG_PX / 4) + j;
                const float4 xv = *reinterpret_cast<const float4*>(&xs[px][ci4 * 4]);
                const float4 dv = *reinterpret_cast<const float4*>(&ds[px][co4 * 4]);
                const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, da[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(xa[i], da[k], acc[i][k]);
                if (want_bias) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) bsum[k] += da[k];
                }
            }
        }
        __syncthreads();
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ci = ci0 + ci4 * 4 + i;
            if (ci >= a.cx) continue;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int co = co4 * 4 + k;
                if (co < a.cy) atomicAdd(a.dw + ((size_t)tap * a.cx + ci) * a.cy + co, acc[i][k]);
            }
        }
        if (want_bias) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (co4 * 4 + k < a.cy) atomicAdd(a.db + co4 * 4 + k, bsum[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ weight gradient, 3x3 stride 1
// The per-plane layers (the five cells, omega's stem): the same sum as conv_wgrad_kernel, organised for FMA issue instead of
// shared-memory traffic.  A CTA stages a tile of 8 x 32 output pixels ONCE for all nine taps -- X with its halo, dY, both zero
// padded outside the image so the inner loop has no predicates -- and each of its nine warps owns one tap: a lane accumulates
// 4 input x 8 output channels, i.e. 32 FMAs per pixel against three LDS.128 (X: one line shared by the lanes of a channel
// quad, dY: two lines of 16 bytes per lane, contiguous over the lanes).  LCO lanes sit along the output channels:
// LCO = 8 / 4 / 2 covers 64 / 32 / 16 output x 16 / 32 / 64 input channels per CTA (blockIdx.y walks the input channel
// blocks).  Two warps per tap (the upper and the lower four rows of the tile: 18 warps spread evenly over the four schedulers);
// a CTA walks `gridDim.x`-strided tiles with its accumulators in registers, the two halves are added through shared memory and
// leave as fp32 atomics once at the end.  Staging is cp.async (all loads of a tile in flight at once).
constexpr int W3_TH = 8, W3_TW = 32, W3_IH = W3_TH + 2, W3_IW = W3_TW + 2, W3_THREADS = 576;

template <int LCO>
__global__ void __launch_bounds__(W3_THREADS) conv_wgrad3_kernel(const __grid_constant__ WGArgs a) {
    constexpr int CO = 8 * LCO, CI = 4 * (32 / LCO);
    extern __shared__ __align__(16) float sm3[];
    float* xs = sm3;                                     // [W3_IH][W3_IW][CI]
    float* ds = sm3 + W3_IH * W3_IW * CI;                // [W3_TH * W3_TW][CO]
    const int tid = threadIdx.x, lane = tid & 31, tap = (tid >> 5) % 9, half = tid / 288;
    const int ky = tap / 3, kx = tap - ky * 3;
    const int ci4 = lane / LCO, co8 = lane % LCO;
    const int ci0 = blockIdx.y * CI;
    const int tiles_x = (a.Wy + W3_TW - 1) / W3_TW, tiles_y = (a.Hy + W3_TH - 1) / W3_TH;
    const bool bias_warp = a.db && tap == 0 && blockIdx.y == 0;          // warp uniform
    const bool want_bias = bias_warp && ci4 == 0;
    float acc[4][8] = {};
    float bsum[8] = {};
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int oy0 = ty * W3_TH, ox0 = tx * W3_TW;
        for (int it = tid; it < W3_IH * W3_IW * (CI / 4); it += W3_THREADS) {
            const int pos = it / (CI / 4), c4 = it - pos * (CI / 4);
            const int r = pos / W3_IW, cc = pos - r * W3_IW;
            const int y = oy0 - 1 + r, x = ox0 - 1 + cc, c = ci0 + c4 * 4;
            const bool ok = y >= 0 && y < a.Hx && x >= 0 && x < a.Wx && c < a.cx;
            const float* gp = a.dy;
            if (ok) {
                int base = 0, sg = 0;
                while (sg < a.nx - 1 && c >= base + a.x[sg].ch) { base += a.x[sg].ch; ++sg; }
                gp = a.x[sg].ptr + ((size_t)y * a.Wx + x) * a.x[sg].ch + (c - base);
            }
            cp_async16_zfill(xs + pos * CI + c4 * 4, gp, ok);
        }
        for (int it = tid; it < W3_TH * W3_TW * (CO / 4); it += W3_THREADS) {
            const int px = it / (CO / 4), q = it - px * (CO / 4);
            const int y = oy0 + px / W3_TW, x = ox0 + px % W3_TW;
            const bool ok = y < a.Hy && x < a.Wy && q * 4 < a.cy;
            cp_async16_zfill(ds + px * CO + q * 4, ok ? a.dy + ((size_t)y * a.Wy + x) * a.cy + q * 4 : a.dy, ok);
        }
        cp_async_wait_all();
        __syncthreads();
        const float* xt = xs + (ky * W3_IW + kx) * CI + ci4 * 4;
        const float* dt = ds + co8 * 4;
        auto sweep = [&](auto with_bias) {               // (the bias sum costs 8 issue slots per pixel: only the warps that own it)
#pragma unroll 1
            for (int r = half * (W3_TH / 2); r < (half + 1) * (W3_TH / 2); ++r) {
#pragma unroll 4
                for (int c = 0; c < W3_TW; ++c) {
                    const float4 xv = ld4(xt + (r * W3_IW + c) * CI);
                    const float4 d0 = ld4(dt + (r * W3_TW + c) * CO), d1 = ld4(dt + (r * W3_TW + c) * CO + CO / 2);
                    const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
                    const float da[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc[i][k] = fmaf(xa[i], da[k], acc[i][k]);
                    if constexpr (decltype(with_bias)::value) {
#pragma unroll
                        for (int k = 0; k < 8; ++k) bsum[k] += da[k];
                    }
                }
            }
        };
        if (bias_warp) sweep(std::true_type{});
        else sweep(std::false_type{});
        __syncthreads();
    }
    // the lower half's sums join the upper half's through shared memory ([accumulator][tap][lane]: conflict free)
    if (half == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) sm3[(i * 8 + k) * 288 + tap * 32 + lane] = acc[i][k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int co = (k < 4 ? 0 : CO / 2) + co8 * 4 + (k & 3);
        if (co >= a.cy) continue;
        if (half == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ci = ci0 + ci4 * 4 + i;
                if (ci < a.cx)
                    atomicAdd(a.dw + ((size_t)tap * a.cx + ci) * a.cy + co, acc[i][k] + sm3[(i * 8 + k) * 288 + tap * 32 + lane]);
            }
        }
        if (want_bias) atomicAdd(a.db + co, bsum[k]);
    }
}

template <int LCO>
static int launch_wgrad3(const WGArgs& a, cudaStream_t stream) {
    constexpr int CO = 8 * LCO, CI = 4 * (32 / LCO);
    constexpr int smem = (W3_IH * W3_IW * CI + W3_TH * W3_TW * CO) * (int)sizeof(float);
    static_assert(smem >= 32 * 288 * (int)sizeof(float), "the tile buffers double as the reduction scratch");
    static unsigned long long smem_ok = 0;
    if (cudaError_t e = ensure_dyn_smem(conv_wgrad3_kernel<LCO>, smem, smem_ok)) return (int)e;
    const int tiles = ((a.Wy + W3_TW - 1) / W3_TW) * ((a.Hy + W3_TH - 1) / W3_TH);
    const int blocks_y = (a.cx + CI - 1) / CI;
    // two tiles per CTA on the large maps: half the atomics, still one wave of the machine across the concurrent samples
    int gx = tiles;
    if (tiles * blocks_y > 148) gx = (tiles + 1) / 2;
    conv_wgrad3_kernel<LCO><<<dim3(gx, blocks_y), W3_THREADS, smem, stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}

// The same for a layer with 1 or 4 output channels (omega's stem conv: 32 -> 4; the head: 8 -> 1): lane = input channel, CY
// accumulators per lane, one warp per tap.  The work is a hundredth of a cell's; the tiled kernel above would spend a full
// 16 x 64 channel block on it.
template <int CY>
__global__ void __launch_bounds__(288) conv_wgrad3_narrow_kernel(const __grid_constant__ WGArgs a) {
    __shared__ __align__(16) float xs[W3_IH * W3_IW * 32];
    __shared__ __align__(16) float ds[W3_TH * W3_TW * CY];
    const int tid = threadIdx.x, lane = tid & 31, tap = tid >> 5;
    const int ky = tap / 3, kx = tap - ky * 3;
    const int tiles_x = (a.Wy + W3_TW - 1) / W3_TW, tiles_y = (a.Hy + W3_TH - 1) / W3_TH;
    const bool bias_warp = a.db && tap == 0;
    float acc[CY] = {}, bsum[CY] = {};
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int oy0 = ty * W3_TH, ox0 = tx * W3_TW;
        for (int it = tid; it < W3_IH * W3_IW * 8; it += 288) {
            const int pos = it >> 3, c = (it & 7) * 4;
            const int r = pos / W3_IW, cc = pos - r * W3_IW;
            const int y = oy0 - 1 + r, x = ox0 - 1 + cc;
            const bool ok = y >= 0 && y < a.Hx && x >= 0 && x < a.Wx && c < a.cx;
            const float* gp = a.dy;
            if (ok) {
                int base = 0, sg = 0;
                while (sg < a.nx - 1 && c >= base + a.x[sg].ch) { base += a.x[sg].ch; ++sg; }
                gp = a.x[sg].ptr + ((size_t)y * a.Wx + x) * a.x[sg].ch + (c - base);
            }
            cp_async16_zfill(xs + pos * 32 + c, gp, ok);
        }
        for (int px = tid; px < W3_TH * W3_TW; px += 288) {
            const int y = oy0 + px / W3_TW, x = ox0 + px % W3_TW;
            const bool ok = y < a.Hy && x < a.Wy;
            const float* dp = a.dy + ((size_t)y * a.Wy + x) * CY;
            if constexpr (CY == 4) cp_async16_zfill(ds + px * 4, ok ? dp : a.dy, ok);
            else ds[px] = ok ? *dp : 0.f;
        }
        cp_async_wait_all();
        __syncthreads();
        const float* xt = xs + (ky * W3_IW + kx) * 32 + lane;
#pragma unroll 1
        for (int r = 0; r < W3_TH; ++r) {
#pragma unroll 8
            for (int c = 0; c < W3_TW; ++c) {
                const float xv = xt[(r * W3_IW + c) * 32];
                float da[CY];
                if constexpr (CY == 4) {
                    const float4 d = ld4(ds + (r * W3_TW + c) * 4);
                    da[0] = d.x; da[1] = d.y; da[2] = d.z; da[3] = d.w;
                } else {
                    da[0] = ds[r * W3_TW + c];
                }
#pragma unroll
                for (int k = 0; k < CY; ++k) acc[k] = fmaf(xv, da[k], acc[k]);
                if (bias_warp) {
#pragma unroll
                    for (int k = 0; k < CY; ++k) bsum[k] += da[k];
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < CY; ++k) {
        if (lane < a.cx) atomicAdd(a.dw + ((size_t)tap * a.cx + lane) * CY + k, acc[k]);
        if (bias_warp && lane == 0) atomicAdd(a.db + k, bsum[k]);
    }
}

template <int CY>
static int launch_wgrad3_narrow(const WGArgs& a, cudaStream_t stream) {
    const int tiles = ((a.Wy + W3_TW - 1) / W3_TW) * ((a.Hy + W3_TH - 1) / W3_TH);
    conv_wgrad3_narrow_kernel<CY><<<tiles > 32 ? (tiles + 1) / 2 : tiles, 288, 0, stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}

// ------------------------------------------------------------------------------------------------ data gradient, 3x3 stride 1
// dX = dG (*) W^T of the five cells (the widest convolutions of the reverse sweep), register tiled like the forward's
// conv_kernel (conv.cu): weights [tap][CIN][COUT] and an input tile + halo in shared memory, lane = tile column, a warp owns
// TCO output channels and TPX tile rows per lane (TPX * TCO accumulators, weights are warp-broadcast LDS.128).
// Out[Y][X][co] = sum_{ky,kx,ci} In[Y + 1 - ky][X + 1 - kx][ci] * w[ky * 3 + kx][ci][co]  (conv_nhwc_kernel's `transposed`
// indexing at stride 1); the output channels are scattered to up to three tensors, stored or accumulated.
struct DXArgs {
    const float* in;       // [H][W][CIN]
    Seg out[3];
    int nout, H, W;
    const float* w;        // [9][CIN][COUT]
};

template <int CIN, int COUT, int TCO, int TPX>
struct DXCfg {
    static constexpr int NWARPS = 8, NCG = COUT / TCO, PSETS = NWARPS / NCG;
    static constexpr int TH = PSETS * TPX, TW = 32, IH = TH + 2, IW = TW + 2;
    static constexpr int CPAD = ((CIN / 4) & 1) ? CIN : CIN + 4;     // odd number of 16-byte chunks: conflict-free LDS.128
    static constexpr int W_FLOATS = 9 * CIN * COUT, IN_FLOATS = IH * IW * CPAD;
    static constexpr int SMEM = (W_FLOATS + IN_FLOATS) * (int)sizeof(float);
    static_assert(COUT % TCO == 0 && NWARPS % NCG == 0 && TCO % 4 == 0 && CIN % 4 == 0, "bad tiling");
    static_assert(SMEM <= 227 * 1024, "tile does not fit");
};

template <int CIN, int COUT, int TCO, int TPX>
__global__ void __launch_bounds__(256, 1) conv_dx3_kernel(const __grid_constant__ DXArgs a) {
    using C = DXCfg<CIN, COUT, TCO, TPX>;
    extern __shared__ __align__(16) float smx[];
    float* sw = smx;
    float* sin = smx + C::W_FLOATS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cg = warp % C::NCG, ps = warp / C::NCG;
    const int co0 = cg * TCO;
    for (int i = tid * 4; i < C::W_FLOATS; i += 256 * 4) cp_async16_zfill(sw + i, a.w + i, true);     // lands with the first tile
    const int tiles_x = (a.W + C::TW - 1) / C::TW, tiles_y = (a.H + C::TH - 1) / C::TH;
    constexpr int CIN4 = CIN / 4;
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int oy0 = ty * C::TH, ox0 = tx * C::TW;
        for (int it = tid; it < C::IH * C::IW * CIN4; it += 256) {
            const int pos = it / CIN4, c4 = it - pos * CIN4;
            const int r = pos / C::IW, cc = pos - r * C::IW;
            const int y = oy0 - 1 + r, x = ox0 - 1 + cc;
            const bool ok = y >= 0 && y < a.H && x >= 0 && x < a.W;
            cp_async16_zfill(sin + pos * C::CPAD + c4 * 4, ok ? a.in + ((size_t)y * a.W + x) * CIN + c4 * 4 : a.in, ok);
        }
        cp_async_wait_all();
        __syncthreads();
        float acc[TPX][TCO];
#pragma unroll
        for (int p = 0; p < TPX; ++p)
#pragma unroll
            for (int k = 0; k < TCO; ++k) acc[p][k] = 0.f;
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap) {
            const int ky = tap / 3, kx = tap - ky * 3;
            const float* wt = sw + tap * CIN * COUT + co0;
            const float* it = sin + ((ps * TPX + 2 - ky) * C::IW + lane + 2 - kx) * C::CPAD;
#pragma unroll 2
            for (int c4 = 0; c4 < CIN4; ++c4) {
                float4 x[TPX];
#pragma unroll
                for (int p = 0; p < TPX; ++p) x[p] = ld4(it + p * C::IW * C::CPAD + c4 * 4);
#pragma unroll
                for (int ci = 0; ci < 4; ++ci) {
                    float wv[TCO];
#pragma unroll
                    for (int k = 0; k < TCO; k += 4) {
                        const float4 w4 = ld4(wt + (c4 * 4 + ci) * COUT + k);
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
        const int ox = ox0 + lane;
#pragma unroll
        for (int p = 0; p < TPX; ++p) {
            const int oy = oy0 + ps * TPX + p;
            if (oy >= a.H || ox >= a.W) continue;
            const size_t pix = (size_t)oy * a.W + ox;
#pragma unroll
            for (int k = 0; k < TCO; k += 4) {
                const int co = co0 + k;
                int base = 0, sg = 0;
                while (sg < a.nout - 1 && co >= base + a.out[sg].ch) { base += a.out[sg].ch; ++sg; }
                float* q = a.out[sg].ptr + pix * a.out[sg].ch + (co - base);
                float4 o = make_float4(acc[p][k], acc[p][k + 1], acc[p][k + 2], acc[p][k + 3]);
                if (a.out[sg].acc) {
                    const float4 t = ld4(q);
                    o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
                }
                st4(q, o);
            }
        }
        __syncthreads();
    }
}

template <int CIN, int COUT, int TCO, int TPX>
static int launch_dx3(const DXArgs& a, cudaStream_t stream) {
    using C = DXCfg<CIN, COUT, TCO, TPX>;
    auto kern = conv_dx3_kernel<CIN, COUT, TCO, TPX>;
    static unsigned long long smem_ok = 0;
    if (cudaError_t e = ensure_dyn_smem(kern, C::SMEM, smem_ok)) return (int)e;
    const int tiles = ((a.W + C::TW - 1) / C::TW) * ((a.H + C::TH - 1) / C::TH);
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    kern<<<tiles < sms ? tiles : sms, 256, C::SMEM, stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}

// ------------------------------------------------------------------------------------------------ GroupNorm
// x [P][C]; group g = channels [g * C / groups, (g + 1) * C / groups) over all P pixels; sums = fp64 (sum, sum of squares).
// 256 % C == 0 (every layer of the path): a thread keeps its channel over the grid-stride loop, partial sums are combined per
// channel through shared memory (no shared-memory fp64 atomics: those are CAS loops) and leave as 2 * groups atomics per block.
__global__ void __launch_bounds__(256) gn_sums_kernel(const float* __restrict__ x, size_t n /* P * C */, int C, int groups,
                                                      double* __restrict__ sums) {
    __shared__ double rs[256], rq[256];
    const int gch = C / groups;
    double s = 0.0, q = 0.0;
    float fs = 0.f, fq = 0.f;
    int it = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        fs += v;
        fq = fmaf(v, v, fq);
        if (++it == 8) { s += fs; q += fq; fs = fq = 0.f; it = 0; }
    }
    rs[threadIdx.x] = s + fs;
    rq[threadIdx.x] = q + fq;
    __syncthreads();
    if ((int)threadIdx.x < C) {                      // channel totals of the block (thread j held channel j % C)
        double a = 0.0, b = 0.0;
        for (int j = threadIdx.x; j < 256; j += C) { a += rs[j]; b += rq[j]; }
        rs[threadIdx.x] = a;
        rq[threadIdx.x] = b;
    }
    __syncthreads();
    if ((int)threadIdx.x < groups) {
        double a = 0.0, b = 0.0;
        for (int c = threadIdx.x * gch; c < (threadIdx.x + 1) * gch; ++c) { a += rs[c]; b += rq[c]; }
        atomicAdd(sums + threadIdx.x * 2, a);
        atomicAdd(sums + threadIdx.x * 2 + 1, b);
    }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x, size_t n, int C, int groups, size_t P,
                                                       const double* __restrict__ sums, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int relu, float* __restrict__ y) {
    __shared__ float sc[64], sf[64];
    if (threadIdx.x < C) {
        const int gch = C / groups;
        float mean, rstd;
        gn_mean_rstd(sums + (threadIdx.x / gch) * 2, (double)gch * (double)P, mean, rstd);
        sc[threadIdx.x] = rstd * gamma[threadIdx.x];
        sf[threadIdx.x] = beta[threadIdx.x] - mean * rstd * gamma[threadIdx.x];
    }
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float v = fmaf(x[i], sc[c], sf[c]);
        y[i] = relu ? fmaxf(v, 0.f) : v;
    }
}

// Backward of y = [ReLU](gamma * xhat + beta), xhat = (x - mean) * rstd over a group of M = (C / groups) * P elements:
//   dz = dy * [y > 0];  S1[c] = sum_p dz * xhat,  S2[c] = sum_p dz            (stage 1: chan_sums [C][2], fp64)
//   A_g = sum_{c in g} gamma_c S1[c],  B_g = sum_{c in g} gamma_c S2[c]
//   dx = rstd * (gamma_c * dz - (xhat * A_g + B_g) / M);  dgamma += S1, dbeta += S2   (stage 2)
struct GnBwdArgs {
    const float* x;
    const float* dy;
    size_t P;
    int C, groups, relu;
    const double* sums;
    const float* gamma;
    const float* beta;
    double* chan_sums;     // [C][2]
    float* dx;             // stage 2 (may alias dy)
    float* dgamma;         // stage 2, +=
    float* dbeta;
    float* dx_chan_sum;    // stage 2, optional [C] += sum_p dx (the bias gradient of the producing conv)
};

template <int STAGE>
__global__ void __launch_bounds__(256) gn_bwd_kernel(const __grid_constant__ GnBwdArgs a) {
    __shared__ float s_mean[8], s_rstd[8], s_A[8], s_B[8];
    __shared__ double r1[256], r2[256];
    const int C = a.C, gch = C / a.groups;
    const double M = (double)gch * (double)a.P;
    if (threadIdx.x < a.groups) {
        float mean, rstd;
        gn_mean_rstd(a.sums + threadIdx.x * 2, M, mean, rstd);
        s_mean[threadIdx.x] = mean;
        s_rstd[threadIdx.x] = rstd;
        if (STAGE == 2) {
            double A = 0.0, B = 0.0;
            for (int c = threadIdx.x * gch; c < (threadIdx.x + 1) * gch; ++c) {
                A += (double)a.gamma[c] * a.chan_sums[c * 2];
                B += (double)a.gamma[c] * a.chan_sums[c * 2 + 1];
            }
            s_A[threadIdx.x] = (float)(A / M);
            s_B[threadIdx.x] = (float)(B / M);
        }
    }
    __syncthreads();
    const size_t n = a.P * (size_t)C;
    // grid stride is a multiple of C (launcher), so a thread always meets the same channel
    const size_t first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (int)(first % C), g = c / gch;
    const float mean = s_mean[g], rstd = s_rstd[g], gam = a.gamma[c], bet = a.beta[c];
    double t1 = 0.0, t2 = 0.0;
    float f1 = 0.f, f2 = 0.f;
    int it = 0;
    for (size_t i = first; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float xh = (a.x[i] - mean) * rstd;
        float dz = a.dy[i];
        if (a.relu && fmaf(gam, xh, bet) <= 0.f) dz = 0.f;
        if (STAGE == 1) {
            f1 += dz * xh;
            f2 += dz;
        } else {
            const float dx = rstd * (gam * dz - (xh * s_A[g] + s_B[g]));
            a.dx[i] = dx;
            f1 += dx;
        }
        if (++it == 16) { t1 += f1; t2 += f2; f1 = f2 = 0.f; it = 0; }
    }
    t1 += f1; t2 += f2;
    // per-channel totals of the block through shared memory (thread j holds channel j % C), then one atomic per channel
    if (STAGE == 1 || a.dx_chan_sum) {
        r1[threadIdx.x] = t1;
        r2[threadIdx.x] = t2;
        __syncthreads();
        if ((int)threadIdx.x < C) {
            double u = 0.0, v = 0.0;
            for (int j = threadIdx.x; j < 256; j += C) { u += r1[j]; v += r2[j]; }
            if (STAGE == 1) {
                atomicAdd(a.chan_sums + threadIdx.x * 2, u);
                atomicAdd(a.chan_sums + threadIdx.x * 2 + 1, v);
            } else {
                atomicAdd(a.dx_chan_sum + threadIdx.x, (float)u);
            }
        }
    }
    if (STAGE == 2 && blockIdx.x == 0 && (int)threadIdx.x < C) {
        // (atomics: the samples of a batch run their backward sweeps on concurrent streams into the same accumulators)
        atomicAdd(a.dgamma + threadIdx.x, (float)a.chan_sums[threadIdx.x * 2]);
        atomicAdd(a.dbeta + threadIdx.x, (float)a.chan_sums[threadIdx.x * 2 + 1]);
    }
}

// ------------------------------------------------------------------------------------------------ LSTM gates
// gates [P][4*Ch] = conv output incl. bias, channel blocks i | f | o | g (module.py:83).
__global__ void __launch_bounds__(256) lstm_gates_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                                             const float* __restrict__ c_new, const float* __restrict__ dh,
                                                             float* __restrict__ dc, float* __restrict__ dgates, size_t P, int Ch) {
    const size_t n = P * (size_t)Ch;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / Ch;
        const int c = (int)(i - p * Ch);
        const float* gp = gates + p * 4 * Ch;
        const float gi = sigmoidf_acc(gp[c]), gf = sigmoidf_acc(gp[Ch + c]), go = sigmoidf_acc(gp[2 * Ch + c]);
        const float gg = tanhf(gp[3 * Ch + c]);
        const float cp = c_prev ? c_prev[i] : 0.f;
        const float th = tanhf(c_new[i]);
        const float dhv = dh[i];
        const float dct = dc[i] + dhv * go * (1.f - th * th);
        float* dg = dgates + p * 4 * Ch;
        dg[c] = dct * gg * gi * (1.f - gi);
        dg[Ch + c] = dct * cp * gf * (1.f - gf);
        dg[2 * Ch + c] = dhv * th * go * (1.f - go);
        dg[3 * Ch + c] = dct * gi * (1.f - gg * gg);
        dc[i] = dct * gf;
    }
}

// The same from the activated gates stored by the training forward (no sigmoid / tanh of pre-activations, no conv recompute).
// max_bits (optional): receives max |dgates| as a bit pattern (atomicMax on the unsigned view of a non-negative float), the
// scale of the tensor-core data gradient that follows (conv_dx_tc.cu); the caller zeroes it.
__global__ void __launch_bounds__(256) lstm_gates_bwd_act_kernel(const float* const* __restrict__ hist, int d_host, const int* __restrict__ d_dev,
                                                                 const float* __restrict__ c_prev, const float* __restrict__ c_new,
                                                                 const float* __restrict__ dh, float* __restrict__ dc,
                                                                 float* __restrict__ dgates, size_t P, int Ch, unsigned* __restrict__ max_bits,
                                                                 float* __restrict__ dbias) {
    __shared__ float red[4][256];
    const float* gates = *hist + (size_t)plane_index(d_host, d_dev) * P * 4 * Ch;
    const size_t n = P * (size_t)Ch;
    float mx = 0.f;
    float sb[4] = {0.f, 0.f, 0.f, 0.f};      // this thread's channel is fixed (256 % Ch == 0): sums of its four gates' gradients
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / Ch;
        const int c = (int)(i - p * Ch);
        const float* gp = gates + p * 4 * Ch;
        const float gi = gp[c], gf = gp[Ch + c], go = gp[2 * Ch + c], gg = gp[3 * Ch + c];
        const float th = tanhf(c_new[i]);
        const float dhv = dh[i];
        const float dct = dc[i] + dhv * go * (1.f - th * th);
        float* dg = dgates + p * 4 * Ch;
        const float d0 = dct * gg * gi * (1.f - gi), d1 = dct * c_prev[i] * gf * (1.f - gf);
        const float d2 = dhv * th * go * (1.f - go), d3 = dct * gi * (1.f - gg * gg);
        dg[c] = d0;
        dg[Ch + c] = d1;
        dg[2 * Ch + c] = d2;
        dg[3 * Ch + c] = d3;
        dc[i] = dct * gf;
        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(d0), fabsf(d1))), fmaxf(fabsf(d2), fabsf(d3)));
        sb[0] += d0; sb[1] += d1; sb[2] += d2; sb[3] += d3;
    }
    if (dbias) {                              // the conv's bias gradient: sum over the pixels (the tensor-core dW has no bias term)
#pragma unroll
        for (int g = 0; g < 4; ++g) red[g][threadIdx.x] = sb[g];
        __syncthreads();
        if ((int)threadIdx.x < 4 * Ch) {
            const int g = threadIdx.x / Ch, c = threadIdx.x - g * Ch;
            float t = 0.f;
            for (int j = c; j < 256; j += Ch) t += red[g][j];
            atomicAdd(dbias + g * Ch + c, t);
        }
    }
    if (max_bits) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(max_bits, __float_as_uint(mx));
    }
}

// ------------------------------------------------------------------------------------------------ max pool 2x2 / 2
__global__ void __launch_bounds__(256) maxpool2_kernel(const float* __restrict__ x, int H, int W, int C, float* __restrict__ out) {
    const int Ho = H / 2, Wo = W / 2;
    const size_t n = (size_t)Ho * Wo * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t pp = i / C;
        const int yo = (int)(pp / Wo), xo = (int)(pp - (size_t)yo * Wo);
        const float* q = x + ((size_t)(2 * yo) * W + 2 * xo) * C + c;
        out[i] = fmaxf(fmaxf(q[0], q[C]), fmaxf(q[(size_t)W * C], q[(size_t)W * C + C]));
    }
}

__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout, int H, int W,
                                                           int C, float* __restrict__ dx) {
    const int Ho = H / 2, Wo = W / 2;
    const size_t n = (size_t)Ho * Wo * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const size_t pp = i / C;
        const int yo = (int)(pp / Wo), xo = (int)(pp - (size_t)yo * Wo);
        const size_t o00 = ((size_t)(2 * yo) * W + 2 * xo) * C + c;
        const size_t off[4] = {o00, o00 + C, o00 + (size_t)W * C, o00 + (size_t)W * C + C};
        int best = 0;
        float bv = x[off[0]];
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            const float v = x[off[k]];
            if (v > bv) { bv = v; best = k; }       // strict: the first maximum in scan order keeps the gradient (ATen)
        }
        dx[off[best]] += dout[i];
    }
}

// reductions that end in per-block atomics: two blocks per SM at most (a few elements per thread instead of one)
static int reduce_grid(size_t n) {
    const size_t b = (n + 255) / 256;
    return (int)(b < 2 * 148 ? (b ? b : 1) : 2 * 148);
}

static int grid_for(size_t n, int threads) {
    size_t b = (n + threads - 1) / threads;
    const size_t cap = 148 * 16;
    return (int)(b < cap ? (b ? b : 1) : cap);
}

}  // namespace

// ================================================================================================ C ABI
extern "C" int aarmvs_conv_nhwc(const aarmvs_seg* in, int nin, int Hin, int Win, const aarmvs_seg* out, int nout, int Hout, int Wout,
                                int cin, int cout, int ksize, int stride, int transposed, const float* weight, const float* bias,
                                void* stream) {
    AARMVS_CHECK_ARG(in && out && weight && nin >= 1 && nin <= 3 && nout >= 1 && nout <= 3);
    AARMVS_CHECK_ARG(Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && (ksize == 1 || ksize == 3) && (stride == 1 || stride == 2));
    GConvArgs a = {};
    int ci = 0, co = 0;
    if (int e = convert_segs(in, nin, a.in, ci)) return e;
    if (int e = convert_segs(out, nout, a.out, co)) return e;
    AARMVS_CHECK_ARG(ci == cin && co == cout);
    a.nin = nin; a.nout = nout; a.Hin = Hin; a.Win = Win; a.Hout = Hout; a.Wout = Wout; a.cin = cin; a.cout = cout;
    a.ks = ksize; a.stride = stride; a.transposed = transposed ? 1 : 0; a.w = weight; a.bias = bias;
    if (transposed && ksize == 3 && stride == 1 && nin == 1 && !a.in[0].acc && !bias && Hin == Hout && Win == Wout) {
        bool quads = true;                               // the cells' data gradients: the register-tiled kernel
        for (int s = 0; s < nout; ++s) quads = quads && a.out[s].ch % 4 == 0;
        if (quads) {
            DXArgs x = {};
            x.in = a.in[0].ptr; x.nout = nout; x.H = Hout; x.W = Wout; x.w = weight;
            for (int s = 0; s < nout; ++s) x.out[s] = a.out[s];
            if (cin == 64 && cout == 48) return launch_dx3<64, 48, 12, 4>(x, (cudaStream_t)stream);      // cells 0, 3
            if (cin == 64 && cout == 32) return launch_dx3<64, 32, 16, 2>(x, (cudaStream_t)stream);      // cells 1, 2
            if (cin == 32 && cout == 40) return launch_dx3<32, 40, 20, 2>(x, (cudaStream_t)stream);      // cell 4
        }
    }
    const int bx = grid_for((size_t)Hout * Wout, GC_THREADS);
    const int co_tile = (bx * ((cout + 15) / 16) < 148 * 3 && cout > 8) ? 8 : 16;      // small maps: more, narrower CTAs
    const int smem = ksize * ksize * cin * co_tile * (int)sizeof(float);
    AARMVS_CHECK_ARG(smem <= 200 * 1024);
    static unsigned long long smem_ok8 = 0, smem_ok16 = 0;
    dim3 grid(bx, (cout + co_tile - 1) / co_tile);
    if (co_tile == 8) {
        if (smem > 48 * 1024)
            if (cudaError_t e = ensure_dyn_smem(conv_nhwc_kernel<8>, 200 * 1024, smem_ok8)) return (int)e;
        conv_nhwc_kernel<8><<<grid, GC_THREADS, smem, (cudaStream_t)stream>>>(a);
    } else {
        if (smem > 48 * 1024)
            if (cudaError_t e = ensure_dyn_smem(conv_nhwc_kernel<16>, 200 * 1024, smem_ok16)) return (int)e;
        conv_nhwc_kernel<16><<<grid, GC_THREADS, smem, (cudaStream_t)stream>>>(a);
    }
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_conv_bwd_weight(const aarmvs_seg* x, int nx, int Hx, int Wx, const float* dy, int Hy, int Wy, int cx, int cy,
                                      int ksize, int stride, float* dweight, float* dbias, void* stream) {
    AARMVS_CHECK_ARG(x && dy && dweight && nx >= 1 && nx <= 3 && Hx > 0 && Wx > 0 && Hy > 0 && Wy > 0);
    AARMVS_CHECK_ARG((ksize == 1 || ksize == 3) && (stride == 1 || stride == 2) && cy >= 1 && cy <= WG_CO);
    WGArgs a = {};
    int c = 0;
    if (int e = convert_segs(x, nx, a.x, c)) return e;
    AARMVS_CHECK_ARG(c == cx);
    a.nx = nx; a.dy = dy; a.Hx = Hx; a.Wx = Wx; a.Hy = Hy; a.Wy = Wy; a.cx = cx; a.cy = cy; a.ks = ksize; a.stride = stride;
    a.dw = dweight; a.db = dbias;
    if (ksize == 3 && stride == 1 && Hx == Hy && Wx == Wy && (cy % 4 == 0 || cy == 1)) {
        bool quads = true;
        for (int s = 0; s < nx; ++s) quads = quads && a.x[s].ch % 4 == 0;
        if (quads && cx <= 32 && cy == 4) return launch_wgrad3_narrow<4>(a, (cudaStream_t)stream);
        if (quads && cx <= 32 && cy == 1) return launch_wgrad3_narrow<1>(a, (cudaStream_t)stream);
        if (quads && cy % 4 == 0) {
            if (cy > 32) return launch_wgrad3<8>(a, (cudaStream_t)stream);
            if (cy > 16) return launch_wgrad3<4>(a, (cudaStream_t)stream);
            return launch_wgrad3<2>(a, (cudaStream_t)stream);
        }
    }
    const int P = Hy * Wy;
    // pixels per block: enough blocks to fill the machine a few times over (the kernel is short), few enough that the fp32
    // atomics at the end stay a small part of it
    const int per_chunk = ksize * ksize * ((cx + WG_CI - 1) / WG_CI);
    int chunks = (P + 4 * WG_PX - 1) / (4 * WG_PX);
    while (chunks > 1 && chunks * per_chunk > 148 * 16) chunks = (chunks + 1) / 2;
    a.chunk = ((P + chunks - 1) / chunks + WG_PX - 1) / WG_PX * WG_PX;
    chunks = (P + a.chunk - 1) / a.chunk;
    dim3 grid(chunks, ksize * ksize, (cx + WG_CI - 1) / WG_CI);
    conv_wgrad_kernel<<<grid, WG_THREADS, 0, (cudaStream_t)stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_gn_sums(const float* x, size_t P, int C, int groups, double* sums, void* stream) {
    AARMVS_CHECK_ARG(x && sums && P > 0 && C >= 1 && C <= 64 && 256 % C == 0 && groups >= 1 && groups <= 8 && C % groups == 0);
    gn_sums_kernel<<<reduce_grid(P * C), 256, 0, (cudaStream_t)stream>>>(x, P * C, C, groups, sums);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_gn_relu_apply(const float* x, size_t P, int C, int groups, const double* sums, const float* gamma,
                                    const float* beta, int relu, float* y, void* stream) {
    AARMVS_CHECK_ARG(x && y && sums && gamma && beta && P > 0 && C >= 1 && C <= 64 && groups >= 1 && groups <= 8 && C % groups == 0);
    gn_apply_kernel<<<grid_for(P * C, 256), 256, 0, (cudaStream_t)stream>>>(x, P * C, C, groups, P, sums, gamma, beta, relu, y);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_gn_relu_bwd(int stage, const float* x, const float* dy, size_t P, int C, int groups, const double* sums,
                                  const float* gamma, const float* beta, int relu, double* chan_sums, float* dx, float* dgamma,
                                  float* dbeta, float* dx_chan_sum, void* stream) {
    AARMVS_CHECK_ARG(x && dy && sums && gamma && beta && chan_sums && P > 0 && (stage == 1 || stage == 2));
    AARMVS_CHECK_ARG(C >= 1 && C <= 64 && 256 % C == 0 && groups >= 1 && groups <= 8 && C % groups == 0);
    AARMVS_CHECK_ARG(stage == 1 || (dx && dgamma && dbeta));
    GnBwdArgs a = {x, dy, P, C, groups, relu, sums, gamma, beta, chan_sums, dx, dgamma, dbeta, dx_chan_sum};
    const int grid = reduce_grid(P * C);            // blockDim (256) is a multiple of C: every thread keeps its channel
    if (stage == 1) gn_bwd_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    else gn_bwd_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_lstm_gates_bwd(const float* gates, const float* c_prev, const float* c_new, const float* dh, float* dc,
                                     float* dgates, size_t P, int hidden, void* stream) {
    AARMVS_CHECK_ARG(gates && c_new && dh && dc && dgates && P > 0 && hidden > 0);
    lstm_gates_bwd_kernel<<<grid_for(P * hidden, 256), 256, 0, (cudaStream_t)stream>>>(gates, c_prev, c_new, dh, dc, dgates, P, hidden);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_lstm_gates_bwd_act(const float* const* gates_history, int d_host, const int* d_dev, const float* c_prev, const float* c_new,
                                         const float* dh, float* dc, float* dgates, size_t P, int hidden, unsigned* max_bits,
                                         float* dbias, void* stream) {
    AARMVS_CHECK_ARG(gates_history && c_prev && c_new && dh && dc && dgates && P > 0 && hidden > 0 && 256 % hidden == 0 && hidden <= 64);
    lstm_gates_bwd_act_kernel<<<grid_for(P * hidden, 256), 256, 0, (cudaStream_t)stream>>>(gates_history, d_host, d_dev, c_prev, c_new, dh,
                                                                                           dc, dgates, P, hidden, max_bits, dbias);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_maxpool2(const float* x, int H, int W, int C, float* out, void* stream) {
    AARMVS_CHECK_ARG(x && out && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0);
    maxpool2_kernel<<<grid_for((size_t)(H / 2) * (W / 2) * C, 256), 256, 0, (cudaStream_t)stream>>>(x, H, W, C, out);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_maxpool2_bwd(const float* x, const float* dout, int H, int W, int C, float* dx, void* stream) {
    AARMVS_CHECK_ARG(x && dout && dx && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C > 0);
    maxpool2_bwd_kernel<<<grid_for((size_t)(H / 2) * (W / 2) * C, 256), 256, 0, (cudaStream_t)stream>>>(x, dout, H, W, C, dx);
    AARMVS_LAUNCH_RESULT();
}
