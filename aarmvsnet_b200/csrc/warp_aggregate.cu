// Plane-sweep warp + inter-view adaptive aggregation.
//
// Reference: models/module.py:6-38 (homo_warping_depthwise), models/drmvsnet.py:27-38 (InterViewAAModule),
// models/drmvsnet.py:307-319 (per-plane aggregation).  The reference materialises, per plane and view, the warped
// volume, its squared difference, every omega activation and three GroupNorm passes.  Here one plane is
//   pass A  cost_stem_kernel   gather S views, (warp - ref)^2 into a shared-memory tile (+1 halo), 3x3 conv 32->4
//                              -> t0 [S][P][4] (16 B/px/view) + GroupNorm-1 sums
//   pass B  omega_stats_kernel<2>   t0 -> GN1/ReLU -> 1x1 -> GroupNorm-2 sums          (reads t0 only)
//   pass C  omega_stats_kernel<3>   ... -> GN2/ReLU -> 1x1 -> GroupNorm-3 sums         (reads t0 only)
//   pass D  aggregate_kernel   re-gather, omega tail from t0 + the three statistics, sum_v (1 + w_v) c_v, * -1/S
// The 5-D cost volume [B,32,D,H,W] never exists, and by default no per-view cost reaches HBM either: pass A keeps only t0
// (16 B per pixel and view) and pass D gathers a second time.  aggregate_stream_kernel below is round 1's alternative -- pass A
// (cost_stem_tc.cu, costv_mode 1) stores the per-view cost of a plane batch (costv [K][S][P][32], 11.6 GB at cfg2) and pass D
// streams it back; as whole jobs on the same box the two are equally fast (the second gather costs what 3 GB/plane of extra
// DRAM traffic costs in clock under the board's power cap: DESIGN.md §3), so the variant without the slab is the default and
// the streaming one stays as Engine(keep_costv=True).  Layout: features NHWC fp32 (one pixel = 128 B = one cache line per
// bilinear tap); 8 lanes x float4 cover one pixel, a warp covers 4 pixels per load.
#include "common.cuh"
#include "tc_common.cuh"
#include "omega.cuh"
#include <string.h>

struct StemParams {            // omega.reweight_network.0.0 repacked [tap][cin][cout] + bias; lives in the
    float w[9 * 32 * 4];       // kernel parameter (constant) bank, so the 1152 FMAs per pixel take their
    float b[4];                // weight operand straight from c[0][...] with no load instruction
};

// ------------------------------------------------------------------------------------------------ plain warp
// The operator seam homo_warping_depthwise(src_fea, src_proj, ref_proj, depth_value): out [P][32].
__global__ void __launch_bounds__(256) warp_kernel(const float* __restrict__ src, float* __restrict__ out, int H, int W,
                                                   Cams cams, int view, float depth, int align_corners) {
    const int chunk = threadIdx.x & 7;
    const int P = H * W;
    for (int p = blockIdx.x * 32 + (threadIdx.x >> 3); p < P; p += gridDim.x * 32) {
        int y = p / W, x = p - y * W;
        float ix, iy;
        sweep_position(cams.m[view], (float)x, (float)y, depth, W, H, align_corners != 0, ix, iy);
        Taps t = make_taps(ix, iy, W, H);
        st4(out + (size_t)p * 32 + chunk * 4, gather_taps(src, t, chunk));
    }
}

// ------------------------------------------------------------------------------------------------ pass A
constexpr int TA_H = 16, TA_W = 32;                     // output tile
constexpr int TA_HH = TA_H + 2, TA_WW = TA_W + 2;       // with the 3x3 halo
constexpr int TA_NPOS = TA_HH * TA_WW;                  // 612 positions
constexpr int TA_THREADS = 128;

// swizzled shared-memory address (floats) of 16-byte chunk j of tile position i: consecutive positions read with
// the same j land in distinct 16-byte bank groups, so both the gather stores and the conv LDS.128 are conflict-free
__device__ __forceinline__ int tile_addr(int i, int j) { return i * 32 + ((j ^ (i & 7)) << 2); }

__global__ void __launch_bounds__(TA_THREADS, 2)
cost_stem_kernel(const float* __restrict__ ref, const float* __restrict__ src, size_t view_stride, int H, int W, int S,
                 const float* __restrict__ cams, const __grid_constant__ StemParams sp,
                 const float* __restrict__ depth_values, int d_host, const int* __restrict__ d_dev, int align_corners,
                 float* __restrict__ t0, double* __restrict__ stats /* [S][3][2] */) {
    extern __shared__ __align__(16) float tile[];        // TA_NPOS * 32 floats
    __shared__ double red[2 * (TA_THREADS / 32)];
    __shared__ float scam[AARMVS_MAX_SRC * 12];
    for (int i = threadIdx.x; i < S * 12; i += TA_THREADS) scam[i] = cams[i];
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = tid & 7, grp = tid >> 3;           // 16 groups of 8 lanes
    const int tx0 = blockIdx.x * TA_W, ty0 = blockIdx.y * TA_H;
    const float depth = depth_values[plane_index(d_host, d_dev) + blockIdx.z];
    const size_t P = (size_t)H * W;
    t0 += (size_t)blockIdx.z * S * P * 4;
    stats += (size_t)blockIdx.z * S * 6;

    for (int v = 0; v < S; ++v) {
        const float* sv = src + v * view_stride;
        // ---- phase 1: (warp - ref)^2 for the tile + halo; zero outside the image (the conv's zero padding)
#pragma unroll 2
        for (int i = grp; i < TA_NPOS; i += TA_THREADS / 8) {
            int r = i / TA_WW, c = i - r * TA_WW;
            int y = ty0 - 1 + r, x = tx0 - 1 + c;
            float4 cv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (y >= 0 && y < H && x >= 0 && x < W) {
                float ix, iy;
                sweep_position(scam + v * 12, (float)x, (float)y, depth, W, H, align_corners != 0, ix, iy);
                Taps t = make_taps(ix, iy, W, H);
                float4 wv = gather_taps(sv, t, chunk);
                float4 rv = ldg4(ref + ((size_t)y * W + x) * 32 + chunk * 4);
                float dx = wv.x - rv.x, dy = wv.y - rv.y, dz = wv.z - rv.z, dw = wv.w - rv.w;
                cv = make_float4(dx * dx, dy * dy, dz * dz, dw * dw);
            }
            st4(tile + tile_addr(i, chunk), cv);
        }
        __syncthreads();
        // ---- phase 2: 3x3 conv 32 -> 4 on the tile; this thread: column `lane`, output rows 4*warp .. 4*warp+3.
        // Sliding window over the 6 input rows: row r feeds output rows r (ky=0), r-1 (ky=1) and r-2 (ky=2), so the
        // loop body is identical for every r and all 1152 weights are compile-time offsets into the parameter bank.
        float s = 0.f, ss = 0.f;
        const int x = tx0 + lane;
        float a0[4], a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int r = 0; r < 6; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) a0[k] = sp.b[k];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc) {
                const int i = (warp * 4 + r) * TA_WW + lane + cc;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 a = ld4(tile + tile_addr(i, j));
                    const float* w0 = sp.w + ((0 * 3 + cc) * 32 + j * 4) * 4;
                    const float* w1 = sp.w + ((1 * 3 + cc) * 32 + j * 4) * 4;
                    const float* w2 = sp.w + ((2 * 3 + cc) * 32 + j * 4) * 4;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        a0[k] = fmaf(a.x, w0[k], a0[k]); a0[k] = fmaf(a.y, w0[4 + k], a0[k]);
                        a0[k] = fmaf(a.z, w0[8 + k], a0[k]); a0[k] = fmaf(a.w, w0[12 + k], a0[k]);
                        a1[k] = fmaf(a.x, w1[k], a1[k]); a1[k] = fmaf(a.y, w1[4 + k], a1[k]);
                        a1[k] = fmaf(a.z, w1[8 + k], a1[k]); a1[k] = fmaf(a.w, w1[12 + k], a1[k]);
                        a2[k] = fmaf(a.x, w2[k], a2[k]); a2[k] = fmaf(a.y, w2[4 + k], a2[k]);
                        a2[k] = fmaf(a.z, w2[8 + k], a2[k]); a2[k] = fmaf(a.w, w2[12 + k], a2[k]);
                    }
                }
            }
            if (r >= 2) {      // output row r - 2 is complete: write t0 and the GroupNorm-1 partial sums
                const int y = ty0 + warp * 4 + r - 2;
                if (y < H && x < W) {
                    st4(t0 + ((size_t)v * P + (size_t)y * W + x) * 4, make_float4(a2[0], a2[1], a2[2], a2[3]));
#pragma unroll
                    for (int k = 0; k < 4; ++k) { s += a2[k]; ss += a2[k] * a2[k]; }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { a2[k] = a1[k]; a1[k] = a0[k]; }
        }
        double dv[2] = {(double)s, (double)ss}, tot[2];
        block_sum<2, TA_THREADS / 32>(dv, red, tot);      // its barriers also protect the tile for the next view
        if (tid == 0) {
            atomicAdd(stats + (v * 3 + 0) * 2 + 0, tot[0]);
            atomicAdd(stats + (v * 3 + 0) * 2 + 1, tot[1]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ omega tail (omega.cuh)
// STAGE 2: accumulate sums of u (GroupNorm-2);  STAGE 3: sums of v (GroupNorm-3).  grid = (blocks, S).
template <int STAGE>
__global__ void __launch_bounds__(256) omega_stats_kernel(const float* __restrict__ t0, size_t P, const __grid_constant__ OmegaTail q,
                                                          double* __restrict__ stats) {
    __shared__ double red[2 * 8];
    const int v = blockIdx.y;
    double* st = stats + v * 6;
    const double cnt = 4.0 * (double)P;
    OmegaNorm n;
    gn_mean_rstd(st + 0, cnt, n.m1, n.r1);
    if (STAGE == 3) gn_mean_rstd(st + 2, cnt, n.m2, n.r2);
    const float* tv = t0 + (size_t)v * P * 4;
    float s = 0.f, ss = 0.f;
    double ds = 0.0, dss = 0.0;
    int it = 0;
    // The pass is pure streaming (16 B per pixel) and was bound by one load in flight per thread (4.0 TB/s): the next
    // pixel's t0 is requested before the current one is evaluated (four more registers, same occupancy).
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4 cur = p < P ? ldg4(tv + p * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    while (p < P) {
        const size_t pn = p + stride;
        float4 nxt = cur;
        if (pn < P) nxt = ldg4(tv + pn * 4);
        float a1[4], u[4], vv[4];
        omega_u(q, n, cur, a1, u);
        if (STAGE == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { s += u[k]; ss += u[k] * u[k]; }
        } else {
            omega_v(q, n, u, vv);
#pragma unroll
            for (int k = 0; k < 4; ++k) { s += vv[k]; ss += vv[k] * vv[k]; }
        }
        if (++it == 8) { ds += s; dss += ss; s = ss = 0.f; it = 0; }   // bound the fp32 partial sums
        cur = nxt;
        p = pn;
    }
    double dv[2] = {ds + s, dss + ss}, tot[2];
    block_sum<2, 8>(dv, red, tot);
    if (threadIdx.x == 0) {
        atomicAdd(st + (STAGE - 1) * 2 + 0, tot[0]);
        atomicAdd(st + (STAGE - 1) * 2 + 1, tot[1]);
    }
}

// ------------------------------------------------------------------------------------------------ pass D
// cost[p][c] = -( sum_v (omega_v(p) + 1) * (warp_v(p)[c] - ref(p)[c])^2 ) / S      (drmvsnet.py:311-319)
// optional output for the operator seam: reweight [S][P].
// 8 lanes x float4 = one pixel; a CTA walks 8 x 16 pixel tiles so the two source rows a bilinear footprint touches are
// shared through L1 between vertically adjacent pixels.  Lane j of a pixel's 8-lane group evaluates view j's homography,
// tap set and omega weight once, the group reads them back with shuffles, and the taps of three views (12 lines) are
// requested before the first is used.
constexpr int AG_TH = 8, AG_TW = 16;
__device__ __forceinline__ float4 ldg4b(const unsigned char* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
#ifndef AGG_NBV
#define AGG_NBV 1              // views whose taps are requested before the first is used
#endif
#ifndef AGG_MINB
#define AGG_MINB 4             // resident CTAs per SM the register allocation aims for
#endif

__global__ void __launch_bounds__(256, AGG_MINB)
aggregate_kernel(const float* __restrict__ ref, const float* __restrict__ src, size_t view_stride, int H, int W, int S, int K,
                 const float* __restrict__ cams, const __grid_constant__ OmegaTail q,
                 const float* __restrict__ depth_values, int d_host, const int* __restrict__ d_dev, int align_corners,
                 const float* __restrict__ t0, const double* __restrict__ stats, float* __restrict__ cost,
                 __half* __restrict__ cost_planes, float* __restrict__ reweight_out) {
    __shared__ OmegaNorm norms_all[AARMVS_MAX_PLANES * AARMVS_MAX_SRC];
    __shared__ float scam[AARMVS_MAX_SRC * 12];
    for (int i = threadIdx.x; i < S * 12; i += blockDim.x) scam[i] = cams[i];
    const size_t P = (size_t)H * W;
    if (threadIdx.x < S * K) {
        OmegaNorm n;
        const double cnt = 4.0 * (double)P;
        gn_mean_rstd(stats + threadIdx.x * 6 + 0, cnt, n.m1, n.r1);
        gn_mean_rstd(stats + threadIdx.x * 6 + 2, cnt, n.m2, n.r2);
        gn_mean_rstd(stats + threadIdx.x * 6 + 4, cnt, n.m3, n.r3);
        norms_all[threadIdx.x] = n;
    }
    __syncthreads();
    const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;
    const float* depths = depth_values + plane_index(d_host, d_dev);
    const float s_den = (float)S;
    const unsigned row_bytes = (unsigned)W * 128u;
    const int tiles_x = (W + AG_TW - 1) / AG_TW, tiles_y = (H + AG_TH - 1) / AG_TH;
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int pl = j * 32 + grp;
            const int y = ty * AG_TH + pl / AG_TW, x = tx * AG_TW + pl % AG_TW;
            const bool valid = y < H && x < W;
            const size_t p = valid ? (size_t)y * W + x : 0;
            const float4 rv = ldg4(ref + p * 32 + sub * 4);
            const uint64_t rv01 = pk2(rv.x, rv.y), rv23 = pk2(rv.z, rv.w);
            // planes innermost: the source lines this pixel touches at plane k+1 are the ones just fetched for plane k
            for (int k = 0; k < K; ++k) {
            const float depth = depths[k];
            const OmegaNorm* norms = norms_all + k * S;
            const float* t0k = t0 + (size_t)k * S * P * 4;
            uint64_t acc01 = 0ull, acc23 = 0ull;
            for (int v0 = 0; v0 < S; v0 += 8) {
                // this lane's view: gather recipe (four weights, 0 for a corner outside the image = the skipped tap; byte offset of
                // the north-west line, as in pass A) and reweight
                const int mv = v0 + sub;
                float4 rw = make_float4(0.f, 0.f, 0.f, 0.f);
                unsigned rb = 0u;
                float wm = 0.f;
                if (mv < S && valid) {
                    make_recipe(scam + mv * 12, x, y, depth, W, H, align_corners != 0, rw, rb);
                    float a1[4], u[4], vv[4];
                    omega_u(q, norms[mv], ldg4(t0k + ((size_t)mv * P + p) * 4), a1, u);
                    omega_v(q, norms[mv], u, vv);
                    wm = omega_w(q, norms[mv], a1, vv);
                    if (reweight_out) reweight_out[((size_t)k * S + mv) * P + p] = wm;
                }
                const int nv = min(8, S - v0);
                const unsigned char* sv = reinterpret_cast<const unsigned char*>(src + (size_t)v0 * view_stride) + sub * 16;
                for (int e = 0; e < nv; ++e, sv += view_stride * sizeof(float)) {
                    // six shuffles hand view e's recipe round the pixel's 8-lane group
                    const unsigned base = __shfl_sync(0xffffffffu, rb, e, 8);
                    const float w0 = __shfl_sync(0xffffffffu, rw.x, e, 8), w1 = __shfl_sync(0xffffffffu, rw.y, e, 8);
                    const float w2 = __shfl_sync(0xffffffffu, rw.z, e, 8), w3 = __shfl_sync(0xffffffffu, rw.w, e, 8);
                    const float wv = __shfl_sync(0xffffffffu, wm, e, 8) + 1.0f;
                    const unsigned char* nl = sv + base;                 // north-west line; the others at fixed steps (make_recipe)
                    const unsigned char* sl = nl + row_bytes;
                    const float4 ta = ldg4b(nl), tb = ldg4b(nl + 128), tcn = ldg4b(sl), td = ldg4b(sl + 128);
                    // blend in ATen's order fma(se, w, fma(sw, w, fma(ne, w, nw * w))), (warp - ref)^2, * (1 + w_v): packed fp32 pairs,
                    // IEEE per lane = bit-identical to the scalar chains
                    const uint64_t p0 = pk2(w0, w0), p1 = pk2(w1, w1), p2 = pk2(w2, w2), p3 = pk2(w3, w3), pv = pk2(wv, wv);
                    const uint64_t g01 = fma2(pk2(td.x, td.y), p3, fma2(pk2(tcn.x, tcn.y), p2, fma2(pk2(tb.x, tb.y), p1, mul2(pk2(ta.x, ta.y), p0))));
                    const uint64_t g23 = fma2(pk2(td.z, td.w), p3, fma2(pk2(tcn.z, tcn.w), p2, fma2(pk2(tb.z, tb.w), p1, mul2(pk2(ta.z, ta.w), p0))));
                    const uint64_t d01 = sub2(g01, rv01), d23 = sub2(g23, rv23);
                    const uint64_t t01 = mul2(pv, mul2(d01, d01)), t23 = mul2(pv, mul2(d23, d23));
                    if (v0 + e == 0) { acc01 = t01; acc23 = t23; }
                    else { acc01 = add2(acc01, t01); acc23 = add2(acc23, t23); }    // the reference adds the rounded product (two ATen ops)
                }
            }
            float4 acc;
            upk2(acc01, acc.x, acc.y);
            upk2(acc23, acc.z, acc.w);
            const float4 r = make_float4(-__fdiv_rn(acc.x, s_den), -__fdiv_rn(acc.y, s_den), -__fdiv_rn(acc.z, s_den), -__fdiv_rn(acc.w, s_den));
            if (cost_planes) {       // fp16 (hi, lo) operand planes [K][2][4 chunks][P][8] of ConvLSTM cell 0, as aggregate_stream_kernel
                uint2 hi, lo;
                tc::split4_f16(r, hi, lo);
                const bool odd = sub & 1;
                const uint2 give = odd ? hi : lo;
                uint2 got;
                got.x = __shfl_xor_sync(0xffffffffu, give.x, 1);
                got.y = __shfl_xor_sync(0xffffffffu, give.y, 1);
                if (valid) {
                    const size_t off = (((size_t)k * 8 + (odd ? 4 : 0) + (sub >> 1)) * P + p) * 8;
                    *reinterpret_cast<uint4*>(cost_planes + off) = odd ? make_uint4(got.x, got.y, lo.x, lo.y) : make_uint4(hi.x, hi.y, got.x, got.y);
                }
            } else if (valid) {
                st4(cost + ((size_t)k * P + p) * 32 + sub * 4, r);
            }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ pass A1: warp cost
// costv[k][v][p][c] = (warp_v(p; depth k)[c] - ref(p)[c])^2 for every pixel, view and plane of the batch, written once;
// the omega stem conv (cost_stem_tc_kernel in stream mode) and pass D then read it back as plain streams, so the gather --
// the expensive, L1-bound part of the plane sweep -- runs once per pixel instead of 2.4 times (halo + second pass).
__global__ void __launch_bounds__(256, 4)
warp_cost_kernel(const float* __restrict__ ref, const float* __restrict__ src, size_t view_stride, int H, int W, int S, int K,
                 const float* __restrict__ cams, const float* __restrict__ depth_values, int d_host,
                 const int* __restrict__ d_dev, int align_corners, float* __restrict__ costv) {
    __shared__ float scam[AARMVS_MAX_SRC * 12];
    for (int i = threadIdx.x; i < S * 12; i += blockDim.x) scam[i] = cams[i];
    __syncthreads();
    const size_t P = (size_t)H * W;
    const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;
    const float* depths = depth_values + plane_index(d_host, d_dev);
    const int tiles_x = (W + AG_TW - 1) / AG_TW, tiles_y = (H + AG_TH - 1) / AG_TH;
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const int pl = j * 32 + grp;
            const int y = ty * AG_TH + pl / AG_TW, x = tx * AG_TW + pl % AG_TW;
            const bool valid = y < H && x < W;
            const size_t p = valid ? (size_t)y * W + x : 0;
            const float4 rv = ldg4(ref + p * 32 + sub * 4);
            for (int k = 0; k < K; ++k) {          // planes innermost: plane k+1 touches the lines just fetched for plane k
                const float depth = depths[k];
                for (int v0 = 0; v0 < S; v0 += 8) {
                    const int mv = v0 + sub;       // lane j of the pixel's 8-lane group evaluates view v0 + j's tap set
                    Taps tm;
                    tm.o_nw = tm.o_ne = tm.o_sw = tm.o_se = -1;
                    tm.w_nw = tm.w_ne = tm.w_sw = tm.w_se = 0.f;
                    if (mv < S && valid) {
                        float ix, iy;
                        sweep_position(scam + mv * 12, (float)x, (float)y, depth, W, H, align_corners != 0, ix, iy);
                        tm = make_taps(ix, iy, W, H);
                    }
                    const int nv = min(8, S - v0);
                    for (int e = 0; e < nv; ++e) {
                        const int o0 = __shfl_sync(0xffffffffu, tm.o_nw, e, 8), o1 = __shfl_sync(0xffffffffu, tm.o_ne, e, 8);
                        const int o2 = __shfl_sync(0xffffffffu, tm.o_sw, e, 8), o3 = __shfl_sync(0xffffffffu, tm.o_se, e, 8);
                        float w0 = __shfl_sync(0xffffffffu, tm.w_nw, e, 8), w1 = __shfl_sync(0xffffffffu, tm.w_ne, e, 8);
                        float w2 = __shfl_sync(0xffffffffu, tm.w_sw, e, 8), w3 = __shfl_sync(0xffffffffu, tm.w_se, e, 8);
                        if (o0 < 0) w0 = 0.f;
                        if (o1 < 0) w1 = 0.f;
                        if (o2 < 0) w2 = 0.f;
                        if (o3 < 0) w3 = 0.f;
                        const float* sv = src + (size_t)(v0 + e) * view_stride + sub * 4;
                        const float4 ta = ldg4(sv + (size_t)max(o0, 0) * 32), tb = ldg4(sv + (size_t)max(o1, 0) * 32);
                        const float4 tcn = ldg4(sv + (size_t)max(o2, 0) * 32), td = ldg4(sv + (size_t)max(o3, 0) * 32);
                        const float gx = __fmaf_rn(td.x, w3, __fmaf_rn(tcn.x, w2, __fmaf_rn(tb.x, w1, __fmul_rn(ta.x, w0))));
                        const float gy = __fmaf_rn(td.y, w3, __fmaf_rn(tcn.y, w2, __fmaf_rn(tb.y, w1, __fmul_rn(ta.y, w0))));
                        const float gz = __fmaf_rn(td.z, w3, __fmaf_rn(tcn.z, w2, __fmaf_rn(tb.z, w1, __fmul_rn(ta.z, w0))));
                        const float gw = __fmaf_rn(td.w, w3, __fmaf_rn(tcn.w, w2, __fmaf_rn(tb.w, w1, __fmul_rn(ta.w, w0))));
                        const float dx = gx - rv.x, dy = gy - rv.y, dz = gz - rv.z, dw = gw - rv.w;
                        if (valid)
                            st4(costv + (((size_t)k * S + v0 + e) * P + p) * 32 + sub * 4,
                                make_float4(__fmul_rn(dx, dx), __fmul_rn(dy, dy), __fmul_rn(dz, dz), __fmul_rn(dw, dw)));
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ pass D, streamed
// Same result as aggregate_kernel, but (warp - ref)^2 is read back from pass A's costv [K][S][P][32] instead of being
// gathered a second time: a pure HBM stream (S * 128 + S * 16 B read, 128 B written per pixel and plane).
#ifndef AGS_PX
#define AGS_PX 1               // pixels per 8-lane group and loop iteration (bytes in flight)
#endif
#ifndef AGS_MINB
#define AGS_MINB 4
#endif
__global__ void __launch_bounds__(256, AGS_MINB)
aggregate_stream_kernel(const float* __restrict__ costv, int S, int K, size_t P, const __grid_constant__ OmegaTail q,
                        const float* __restrict__ t0, const double* __restrict__ stats, float* __restrict__ cost,
                        __half* __restrict__ cost_planes, float* __restrict__ reweight_out) {
    __shared__ OmegaNorm norms_all[AARMVS_MAX_PLANES * AARMVS_MAX_SRC];
    if (threadIdx.x < S * K) {
        OmegaNorm n;
        const double cnt = 4.0 * (double)P;
        gn_mean_rstd(stats + threadIdx.x * 6 + 0, cnt, n.m1, n.r1);
        gn_mean_rstd(stats + threadIdx.x * 6 + 2, cnt, n.m2, n.r2);
        gn_mean_rstd(stats + threadIdx.x * 6 + 4, cnt, n.m3, n.r3);
        norms_all[threadIdx.x] = n;
    }
    __syncthreads();
    const int sub = threadIdx.x & 7;
    const float s_den = (float)S;
    const size_t total = P * K, stride = (size_t)gridDim.x * 32;
    for (size_t i0 = (size_t)blockIdx.x * 32 + (threadIdx.x >> 3); i0 < total; i0 += stride * AGS_PX) {
        float4 acc[AGS_PX];
#pragma unroll
        for (int u = 0; u < AGS_PX; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int v0 = 0; v0 < S; v0 += 8) {
            const int nv = min(8, S - v0), mv = v0 + sub;
            float4 c[AGS_PX][8], tq[AGS_PX];
            // all loads of this step first
#pragma unroll
            for (int u = 0; u < AGS_PX; ++u) {
                const size_t i = min(i0 + u * stride, total - 1);
                const int k = (int)(i / P);
                const size_t p = i - (size_t)k * P;
                tq[u] = mv < S ? ldg4(t0 + ((size_t)(k * S + mv) * P + p) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (e < nv) c[u][e] = ldg4(costv + ((size_t)(k * S + v0 + e) * P + p) * 32 + sub * 4);
            }
#pragma unroll
            for (int u = 0; u < AGS_PX; ++u) {
                const size_t i = min(i0 + u * stride, total - 1);
                const int k = (int)(i / P);
                const size_t p = i - (size_t)k * P;
                const OmegaNorm* norms = norms_all + k * S;
                float wm = 0.f;
                if (mv < S) {            // lane j of the pixel's 8-lane group evaluates view v0 + j's reweight
                    float a1[4], uu[4], vv[4];
                    omega_u(q, norms[mv], tq[u], a1, uu);
                    omega_v(q, norms[mv], uu, vv);
                    wm = omega_w(q, norms[mv], a1, vv);
                    if (reweight_out && i0 + u * stride < total) reweight_out[((size_t)k * S + mv) * P + p] = wm;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float wv = __shfl_sync(0xffffffffu, wm, e, 8) + 1.0f;
                    if (e >= nv) continue;
                    const float4 ce = c[u][e];
                    if (v0 + e == 0) {
                        acc[u] = make_float4(__fmul_rn(wv, ce.x), __fmul_rn(wv, ce.y), __fmul_rn(wv, ce.z), __fmul_rn(wv, ce.w));
                    } else {
                        acc[u].x = __fadd_rn(acc[u].x, __fmul_rn(wv, ce.x)); acc[u].y = __fadd_rn(acc[u].y, __fmul_rn(wv, ce.y));
                        acc[u].z = __fadd_rn(acc[u].z, __fmul_rn(wv, ce.z)); acc[u].w = __fadd_rn(acc[u].w, __fmul_rn(wv, ce.w));
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < AGS_PX; ++u) {
            const size_t i = i0 + u * stride;
            const float4 r = make_float4(-__fdiv_rn(acc[u].x, s_den), -__fdiv_rn(acc[u].y, s_den),
                                         -__fdiv_rn(acc[u].z, s_den), -__fdiv_rn(acc[u].w, s_den));
            if (cost_planes) {       // fp16 (hi, lo) operand planes [K][2][4 chunks][P][8] for the TMA-fed ConvLSTM cell 0
                // lanes (sub, sub ^ 1) hold the two halves of one 8-channel chunk: the even lane stores the chunk's 16-byte hi
                // element, the odd lane its lo element (the whole 8-lane group is in or out of range together)
                uint2 hi, lo;
                tc::split4_f16(r, hi, lo);
                const bool odd = sub & 1;
                const uint2 give = odd ? hi : lo;
                uint2 got;
                got.x = __shfl_xor_sync(0xffffffffu, give.x, 1);
                got.y = __shfl_xor_sync(0xffffffffu, give.y, 1);
                if (i < total) {
                    const size_t k = i / P, p = i - k * P;
                    const size_t off = ((k * 8 + (odd ? 4 : 0) + (sub >> 1)) * P + p) * 8;
                    *reinterpret_cast<uint4*>(cost_planes + off) = odd ? make_uint4(got.x, got.y, lo.x, lo.y) : make_uint4(hi.x, hi.y, got.x, got.y);
                }
            } else if (i < total) {
                st4(cost + i * 32 + sub * 4, r);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ C ABI
#include "../../include/aarmvs.h"

static int fill_cams(Cams& c, const float* cams_host, int S) {
    if (!cams_host || S < 1 || S > AARMVS_MAX_SRC) return AARMVS_ERR_ARG;
    for (int v = 0; v < S; ++v)
        for (int k = 0; k < 12; ++k) c.m[v][k] = cams_host[v * 12 + k];
    return AARMVS_OK;
}

extern "C" int aarmvs_warp(const float* src, float* out, int H, int W, const float* cam_host, float depth,
                           int align_corners, void* stream) {
    AARMVS_CHECK_ARG(src && out && H > 0 && W > 0);
    Cams c;
    if (int e = fill_cams(c, cam_host, 1)) return e;
    int P = H * W;
    int blocks = min((P + 31) / 32, 148 * 16);
    warp_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, out, H, W, c, 0, depth, align_corners);
    AARMVS_LAUNCH_RESULT();
}

extern "C" size_t aarmvs_omega_stats_bytes(int S) { return (size_t)S * 6 * sizeof(double); }

extern "C" int aarmvs_cost_stem(const float* ref, const float* src, size_t view_stride, int H, int W, int S,
                                const float* cams, const float* stem_w, const float* stem_b,
                                const float* depth_values, int d_host, const int* d_dev, int planes, int align_corners,
                                float* t0, double* stats, void* stream) {
    AARMVS_CHECK_ARG(ref && src && t0 && stats && stem_w && stem_b && depth_values && cams && H > 0 && W > 0);
    AARMVS_CHECK_ARG(S >= 1 && S <= AARMVS_MAX_SRC && planes >= 1 && planes <= AARMVS_MAX_PLANES);
    StemParams sp;
    for (int i = 0; i < 9 * 32 * 4; ++i) sp.w[i] = stem_w[i];
    for (int i = 0; i < 4; ++i) sp.b[i] = stem_b[i];
    const int smem = TA_NPOS * 32 * sizeof(float);
    static unsigned long long smem_ok = 0;
    if (cudaError_t e = ensure_dyn_smem(cost_stem_kernel, smem, smem_ok)) return (int)e;
    dim3 grid((W + TA_W - 1) / TA_W, (H + TA_H - 1) / TA_H, planes);
    cost_stem_kernel<<<grid, TA_THREADS, smem, (cudaStream_t)stream>>>(ref, src, view_stride, H, W, S, cams, sp, depth_values,
                                                                      d_host, d_dev, align_corners, t0, stats);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_warp_cost(const float* ref, const float* src, size_t view_stride, int H, int W, int S, const float* cams,
                                const float* depth_values, int d_host, const int* d_dev, int planes, int align_corners,
                                float* costv, void* stream) {
    AARMVS_CHECK_ARG(ref && src && cams && depth_values && costv && H > 0 && W > 0);
    AARMVS_CHECK_ARG(S >= 1 && S <= AARMVS_MAX_SRC && planes >= 1 && planes <= AARMVS_MAX_PLANES);
    const int ntiles = ((W + AG_TW - 1) / AG_TW) * ((H + AG_TH - 1) / AG_TH);
    const int blocks = ntiles < 148 * 4 ? ntiles : 148 * 4;
    warp_cost_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ref, src, view_stride, H, W, S, planes, cams, depth_values, d_host,
                                                               d_dev, align_corners, costv);
    AARMVS_LAUNCH_RESULT();
}

static void fill_tail(OmegaTail& q, const float* t) { memcpy(&q, t, sizeof(OmegaTail)); }

extern "C" int aarmvs_omega_tail_floats(void) { return (int)(sizeof(OmegaTail) / sizeof(float)); }

extern "C" int aarmvs_omega_stats(int stage, const float* t0, int H, int W, int S, int planes, const float* tail_host,
                                  double* stats, void* stream) {
    AARMVS_CHECK_ARG(t0 && stats && tail_host && (stage == 2 || stage == 3) && S >= 1 && S <= AARMVS_MAX_SRC);
    AARMVS_CHECK_ARG(planes >= 1 && planes <= AARMVS_MAX_PLANES);
    S *= planes;          // t0 is [planes][S][P][4] and stats [planes][S][6]: every (plane, view) pair is one slot
    OmegaTail q;
    fill_tail(q, tail_host);
    size_t P = (size_t)H * W;
    dim3 grid((unsigned)min((P + 255) / 256, (size_t)148 * 4), S);
    if (stage == 2) omega_stats_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(t0, P, q, stats);
    else omega_stats_kernel<3><<<grid, 256, 0, (cudaStream_t)stream>>>(t0, P, q, stats);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_aggregate(const float* ref, const float* src, size_t view_stride, int H, int W, int S,
                                const float* cams, const float* tail_host, const float* depth_values, int d_host,
                                const int* d_dev, int planes, int align_corners, const float* t0, const double* stats,
                                const float* costv, float* cost, void* cost_planes, float* reweight_out, void* stream) {
    AARMVS_CHECK_ARG(t0 && stats && tail_host && H > 0 && W > 0);
    AARMVS_CHECK_ARG(cost_planes ? ((uintptr_t)cost_planes & 15) == 0 : cost != nullptr);
    AARMVS_CHECK_ARG(S >= 1 && S <= AARMVS_MAX_SRC && planes >= 1 && planes <= AARMVS_MAX_PLANES);
    OmegaTail q;
    fill_tail(q, tail_host);
    if (costv) {           // pass A kept (warp - ref)^2: stream it instead of gathering again
        const size_t Pn = (size_t)H * W;
        // the four 8-lane pixel groups of a warp enter and leave the kernel's loop together (its shuffles name the full warp)
        AARMVS_CHECK_ARG(Pn % 4 == 0);
        const size_t groups = (Pn * planes + 31) / 32;
        const int nblk = (int)(groups < (size_t)148 * 32 ? groups : (size_t)148 * 32);
        aggregate_stream_kernel<<<nblk, 256, 0, (cudaStream_t)stream>>>(costv, S, planes, Pn, q, t0, stats, cost, (__half*)cost_planes, reweight_out);
        AARMVS_LAUNCH_RESULT();
    }
    AARMVS_CHECK_ARG(ref && src && depth_values && cams);
    AARMVS_CHECK_ARG((size_t)H * W * 128 < 0xFFFFFFFFull);          // the gather recipes hold 32-bit byte offsets
    // (a column-walking variant organised like pass A's gather -- one thread per pixel for the recipe and the reweight, 8-lane
    //  teams walking eight-pixel column segments with line reuse and register accumulation over the views, bit-identical -- was
    //  measured SLOWER, 0.68 vs 0.57 ms/plane at cfg2: 96-128 registers, one barrier per view and at most 24 resident warps against
    //  this kernel's 32 barrier-free ones; profiles/r02_stage_variants.log)
    const int ntiles = ((W + AG_TW - 1) / AG_TW) * ((H + AG_TH - 1) / AG_TH);
    const int blocks = ntiles < 148 * AGG_MINB ? ntiles : 148 * AGG_MINB;
    aggregate_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ref, src, view_stride, H, W, S, planes, cams, q, depth_values, d_host,
                                                              d_dev, align_corners, t0, stats, cost, (__half*)cost_planes, reweight_out);
    AARMVS_LAUNCH_RESULT();
}
