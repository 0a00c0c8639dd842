// Pass A of the plane sweep on the tensor cores (models/module.py:6-38,
// models/drmvsnet.py:31,311): for every source view, bilinear gather at depth d, c = (warp - ref)^2, and the omega
// stem 3x3 conv 32 -> 4 (omega.reweight_network.0.0) on tcgen05.
//
// The conv is evaluated as  z[p][tap][o] = sum_c c[p][c] * w[o][c][tap]   (one GEMM per staged halo tile: M = the 180
// halo pixels in two 128-row blocks, N = 9 taps x 4 outputs = 36 -> 48, K = 32 channels)  followed by the 9-tap shifted
// sum  t0[q][o] = b[o] + sum_tap z[q + tap][tap][o].  Operands: un-swizzled K-major fp16 (hi, lo) planes
// [8-channel chunk][row][16 B] (see convlstm_tc.cu / tc_common.cuh); fp32 accumulation in TMEM.
//
// Work unit = one "fill" = (tile of 16 x 8 pixels + halo, source view, plane); fills run in (tile, view, plane) order so
// a tile's source lines stay in L1 across the planes of a batch.  The CTA (one per SM, persistent) is a ring of
// independent WORKER GROUPS of four warps plus one MMA-issuing warp.  Group g owns operand stage g, recipe slot g and TMEM
// accumulator g and walks fills g, g+G, g+2G, ...:
//   1. one thread per staged pixel evaluates the homography and writes a ready-to-use gather recipe (shared memory);
//   2. 8 lanes per pixel gather branch-free (out-of-image taps read line 0 with weight 0), (warp-ref)^2, fp16 split,
//      stage -> arrive on full[g];                      [the MMA warp issues the 12 MMAs of that fill, commit -> done[g]]
//   3. wait done[g]: TMEM -> z rows -> exchange through the (now free) operand stage -> 9-tap sum -> t0, GroupNorm-1 sums.
// The groups are out of phase, so one group's load / MMA latency is covered by the others' arithmetic, and no role
// waits on another group's progress (the first version had a single epilogue group that serialised the pipeline).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/aarmvs.h"

namespace {

#ifndef STEM_GROUPS
#define STEM_GROUPS 5
#endif
#ifndef STEM_NB
#define STEM_NB 2
#endif
constexpr int TR = 16, TCOL = 8, HR = 18, HC = 10, NPX = HR * HC;        // 180 staged pixels
#ifndef STEM_AROWS
#define STEM_AROWS 210
#endif
constexpr int AROWS = STEM_AROWS;          // operand rows per K chunk: >= NPX; = 2 mod 8 puts the four chunks a half-warp's
                                           // 8-byte stores touch (2 pixels x 8 lanes) in four disjoint groups of 8 banks; and
                                           // 2*4*AROWS*16 >= the z exchange buffer that reuses the stage
constexpr int G = STEM_GROUPS;             // worker groups = operand stages = TMEM accumulators
constexpr int GROUP_THREADS = 128;
constexpr int N_THREADS = 32 + G * GROUP_THREADS;
constexpr int KC8 = 4, NZ = 48;                                          // K chunks; accumulator columns (36 used)
constexpr int A_PLANE = KC8 * AROWS * 16, A_STAGE = 2 * A_PLANE;         // 13056, 26112 bytes
constexpr int B_PLANE = KC8 * NZ * 16, B_BYTES = 2 * B_PLANE;            // 3072, 6144 bytes
constexpr int TAP_BYTES = NPX * 48;                                      // one TapRec (48 B) per staged pixel
constexpr int Z_BYTES = 9 * NPX * 16;                                    // z exchange [tap][pixel][4 floats] = 25920
static_assert(Z_BYTES <= A_STAGE, "the z exchange buffer reuses the operand stage");
constexpr int SMEM = B_BYTES + G * (A_STAGE + TAP_BYTES) + 1024;
constexpr int TMEM_COLS = 512;                                           // G accumulators x 2 row blocks x 48 columns
static_assert(G * 2 * NZ <= TMEM_COLS, "accumulators exceed TMEM");

struct StemArgs {
    const float* ref;
    const float* src;
    size_t view_stride;
    int H, W, S, K;             // K = planes per launch
    const float* cams;
    const __half* weight;       // [2 planes hi,lo][4 chunks][48 columns (tap*4 + out)][8] fp16
    float bias[4];
    const float* depth_values;
    int d_host;
    const int* d_dev;
    int align_corners;
    float* t0;
    double* stats;
    float* costv;               // optional [K][S][P][32] (warp-ref)^2: written for the tile's interior pixels (mode 1) so pass D
    int costv_mode;             //   need not gather again, or READ instead of gathering (mode 2: aarmvs_warp_cost wrote it)
    unsigned long long* prof;   // debug: clock64 stamps of group 0 of CTA 0, [64 fills][8]
    int dbg;                    // debug ablation flags (aarmvs_debug_flags): 1 skip gather loads, 2 skip recipes, 4 skip MMA
};

// Gather recipe of one staged pixel, ready to use: float offsets of the four corner lines (out-of-image corners point
// at line 0 with weight 0, which reproduces the skipped tap exactly), and the offset of the reference pixel
// (-1: the pixel lies outside the image, its cost is the conv's zero padding).
struct __align__(16) TapRec {
    int o[4];
    float w[4];
    int ref_off, pad0, pad1, pad2;
};

__global__ void __launch_bounds__(N_THREADS, 1) cost_stem_tc_kernel(const __grid_constant__ StemArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;
    unsigned char* sA = smem + B_BYTES;
    unsigned char* sT = sA + G * A_STAGE;
    __shared__ uint64_t full_a[G], done_a[G];
    __shared__ uint32_t tmem_base_s;
    __shared__ float scam[AARMVS_MAX_SRC * 12];
    __shared__ float2 sstat[G * 4][AARMVS_MAX_PLANES * AARMVS_MAX_SRC];      // per worker warp, per (plane, view): no atomics

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* depths = a.depth_values + plane_index(a.d_host, a.d_dev);

    for (int i = tid; i < B_BYTES / 16; i += N_THREADS)
        reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(a.weight) + i);
    if (a.costv_mode != 2)          // stream mode never evaluates a homography
        for (int i = tid; i < a.S * 12; i += N_THREADS) scam[i] = a.cams[i];
    for (int i = tid; i < G * 4 * AARMVS_MAX_PLANES * AARMVS_MAX_SRC; i += N_THREADS) (&sstat[0][0])[i] = make_float2(0.f, 0.f);
    if (tid == 0) {
        for (int s = 0; s < G; ++s) { tc::mbar_init(&full_a[s], GROUP_THREADS); tc::mbar_init(&done_a[s], 1); }
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc<TMEM_COLS>(&tmem_base_s);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const int tiles_x = (a.W + TCOL - 1) / TCOL, tiles_y = (a.H + TR - 1) / TR;
    const int ntiles = tiles_x * tiles_y;
    const size_t P = (size_t)a.H * a.W;
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int SK = a.S * a.K;
    const int total = my_tiles * SK;

    if (warp == 0) {
        // =========================================== MMA issuer ===========================================
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::idesc_f16_f32(128, NZ);
        const uint32_t a_base = tc::smem_u32(sA), b_base = tc::smem_u32(sB);
        constexpr uint32_t A_HI = tc::desc_hi(8 * 16), B_HI = tc::desc_hi(8 * 16);
        const uint32_t b_lo_hi = tc::desc_lo(b_base, NZ * 16), b_lo_lo = tc::desc_lo(b_base + B_PLANE, NZ * 16);
        for (int it = 0; it < total; ++it) {
            const int g = it % G;
            tc::mbar_wait(&full_a[g], (it / G) & 1);
            tc::tc_fence_after();
            const uint32_t a_lo_hi = tc::desc_lo(a_base + g * A_STAGE, AROWS * 16);
            const uint32_t a_lo_lo = tc::desc_lo(a_base + g * A_STAGE + A_PLANE, AROWS * 16);
            if (leader) {
                if (!(a.dbg & 4))
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    const uint32_t dst = tmem_base + (g * 2 + mb) * NZ;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t ak = (uint32_t)((2 * ks) * AROWS + mb * 128);        // 16-byte units
                        const uint32_t bk = (uint32_t)((2 * ks) * NZ);
                        tc::mma_f16_parts(dst, a_lo_lo + ak, A_HI, b_lo_hi + bk, B_HI, idesc, ks != 0);
                        tc::mma_f16_parts(dst, a_lo_hi + ak, A_HI, b_lo_lo + bk, B_HI, idesc, 1);
                        tc::mma_f16_parts(dst, a_lo_hi + ak, A_HI, b_lo_hi + bk, B_HI, idesc, 1);
                    }
                }
                tc::mma_commit(&done_a[g]);        // operand stage free AND accumulator complete
            }
            __syncwarp();
        }
    } else {
        // =========================================== worker groups ===========================================
        const int g = (warp - 1) / 4, lt = tid - 32 - g * GROUP_THREADS;
        const int quad = warp & 3;                                   // TMEM lane quadrant this warp may read
        unsigned char* stage = sA + g * A_STAGE;
        TapRec* taps = reinterpret_cast<TapRec*>(sT + g * TAP_BYTES);
        const uint32_t bar_id = 1 + g;
        constexpr int ITEMS = (NPX * 8 + GROUP_THREADS - 1) / GROUP_THREADS;     // 12 (the last one partial)
        constexpr int NB = STEM_NB;
        const int q0 = ((lt >> 3) * HC + (lt & 7)) * 16;              // output pixel lt of the tile -> halo pixel, tap (0,0)
        for (int it = g; it < total; it += G) {
            const int tile = blockIdx.x + (it / SK) * gridDim.x, v = (it % SK) / a.K, k = it % a.K;
            const float depth = depths[k];
            const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
            const bool pf = a.prof && blockIdx.x == 0 && g == 0 && lt == 0 && it / G < 64;
            unsigned long long* pp = a.prof + (it / G) * 8;
            if (pf) pp[0] = clock64();
            // ---- 1. gather recipes (not needed when (warp-ref)^2 is streamed from costv)
            for (int px = lt; px < NPX && a.costv_mode != 2; px += GROUP_THREADS) {
                const int r = px / HC, cc = px - r * HC;
                const int y = ty * TR - 1 + r, x = tx * TCOL - 1 + cc;
                TapRec rec;
                rec.o[0] = rec.o[1] = rec.o[2] = rec.o[3] = 0;
                rec.w[0] = rec.w[1] = rec.w[2] = rec.w[3] = 0.f;
                rec.ref_off = -1; rec.pad0 = (r >= 1 && r <= TR && cc >= 1 && cc <= TCOL) ? 1 : 0; rec.pad1 = rec.pad2 = 0;   // pad0: interior pixel
                if (y >= 0 && y < a.H && x >= 0 && x < a.W && !(a.dbg & 2)) {
                    float ix, iy;
                    sweep_position(scam + v * 12, (float)x, (float)y, depth, a.W, a.H, a.align_corners != 0, ix, iy);
                    const Taps t = make_taps(ix, iy, a.W, a.H);
                    rec.o[0] = max(t.o_nw, 0) * 32; rec.o[1] = max(t.o_ne, 0) * 32;
                    rec.o[2] = max(t.o_sw, 0) * 32; rec.o[3] = max(t.o_se, 0) * 32;
                    rec.w[0] = t.o_nw >= 0 ? t.w_nw : 0.f; rec.w[1] = t.o_ne >= 0 ? t.w_ne : 0.f;
                    rec.w[2] = t.o_sw >= 0 ? t.w_sw : 0.f; rec.w[3] = t.o_se >= 0 ? t.w_se : 0.f;
                    rec.ref_off = (y * a.W + x) * 32;
                }
                taps[px] = rec;
            }
            // also orders the previous fill's z reads (step 3) before this fill's stage writes
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(GROUP_THREADS) : "memory");
            if (pf) pp[1] = clock64();
            // ---- 2. gather, (warp - ref)^2, split, stage
            const float* sv = a.src + (size_t)v * a.view_stride;
            float* costv_slot = a.costv ? a.costv + ((size_t)(k * a.S + v) * P) * 32 : nullptr;
            if (a.costv_mode == 2) {
                // stream mode: stage the halo tile straight from costv (coalesced 128-byte lines, no gather, no recipes)
#pragma unroll 1
                constexpr int SB = 6;                 // loads in flight per thread: HBM latency, not bandwidth, bounds this loop
                for (int jb = 0; jb < ITEMS; jb += SB) {
                    float4 cvv[SB];
#pragma unroll
                    for (int j = 0; j < SB; ++j) {
                        const int i = lt + (jb + j) * GROUP_THREADS;
                        const int px = min(i, NPX * 8 - 1) >> 3, c4 = i & 7;
                        const int r = px / HC, cc = px - r * HC;
                        const int y = ty * TR - 1 + r, x = tx * TCOL - 1 + cc;
                        cvv[j] = (i < NPX * 8 && y >= 0 && y < a.H && x >= 0 && x < a.W)
                                     ? ldg4(costv_slot + ((size_t)y * a.W + x) * 32 + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < SB; ++j) {
                        const int i = lt + (jb + j) * GROUP_THREADS;
                        const int px = i >> 3, c4 = i & 7;
                        uint2 hi, lo;
                        tc::split4_f16(cvv[j], hi, lo);
                        if (i < NPX * 8) {
                            const size_t off = ((size_t)(c4 >> 1) * AROWS + px) * 16 + (c4 & 1) * 8;
                            *reinterpret_cast<uint2*>(stage + off) = hi;
                            *reinterpret_cast<uint2*>(stage + A_PLANE + off) = lo;
                        }
                    }
                }
            } else
#pragma unroll 1
            for (int jb = 0; jb < ITEMS; jb += NB) {
                float4 ta[NB], tb[NB], tcn[NB], td[NB], rv[NB], wq[NB];
                int roff[NB], interior[NB];
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const int i = min(lt + (jb + j) * GROUP_THREADS, NPX * 8 - 1);
                    const int px = i >> 3, c4 = (i & 7) * 4;
                    const int4 o = *reinterpret_cast<const int4*>(taps[px].o);
                    wq[j] = *reinterpret_cast<const float4*>(taps[px].w);
                    roff[j] = taps[px].ref_off;
                    interior[j] = taps[px].pad0;
                    if (a.dbg & 1) {
                        ta[j] = tb[j] = tcn[j] = td[j] = rv[j] = make_float4(1.f, 2.f, 3.f, 4.f);
                    } else {
                        ta[j] = ldg4(sv + o.x + c4);
                        tb[j] = ldg4(sv + o.y + c4);
                        tcn[j] = ldg4(sv + o.z + c4);
                        td[j] = ldg4(sv + o.w + c4);
                        rv[j] = ldg4(a.ref + max(roff[j], 0) + c4);
                    }
                }
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const int iraw = lt + (jb + j) * GROUP_THREADS;
                    const int px = iraw >> 3, c4 = iraw & 7;
                    // bilinear blend in ATen's order (see gather_taps)
                    const float gx = __fmaf_rn(td[j].x, wq[j].w, __fmaf_rn(tcn[j].x, wq[j].z, __fmaf_rn(tb[j].x, wq[j].y, __fmul_rn(ta[j].x, wq[j].x))));
                    const float gy = __fmaf_rn(td[j].y, wq[j].w, __fmaf_rn(tcn[j].y, wq[j].z, __fmaf_rn(tb[j].y, wq[j].y, __fmul_rn(ta[j].y, wq[j].x))));
                    const float gz = __fmaf_rn(td[j].z, wq[j].w, __fmaf_rn(tcn[j].z, wq[j].z, __fmaf_rn(tb[j].z, wq[j].y, __fmul_rn(ta[j].z, wq[j].x))));
                    const float gw = __fmaf_rn(td[j].w, wq[j].w, __fmaf_rn(tcn[j].w, wq[j].z, __fmaf_rn(tb[j].w, wq[j].y, __fmul_rn(ta[j].w, wq[j].x))));
                    const float dx = gx - rv[j].x, dy = gy - rv[j].y, dz = gz - rv[j].z, dw = gw - rv[j].w;
                    const bool in = roff[j] >= 0;
                    const float4 cv = make_float4(in ? dx * dx : 0.f, in ? dy * dy : 0.f, in ? dz * dz : 0.f, in ? dw * dw : 0.f);
                    if (a.costv_mode == 1 && in && interior[j] && iraw < NPX * 8)          // interior pixel of this tile, inside the image
                        st4(costv_slot + roff[j] + c4 * 4, cv);
                    uint2 hi, lo;
                    tc::split4_f16(cv, hi, lo);
                    if (iraw < NPX * 8) {
                        const size_t off = ((size_t)(c4 >> 1) * AROWS + px) * 16 + (c4 & 1) * 8;
                        *reinterpret_cast<uint2*>(stage + off) = hi;
                        *reinterpret_cast<uint2*>(stage + A_PLANE + off) = lo;
                    }
                }
            }
            if (pf) pp[2] = clock64();
            tc::fence_async_smem();
            tc::mbar_arrive(&full_a[g]);
            if (pf) pp[3] = clock64();
            // ---- 3. accumulator -> z rows -> exchange through the operand stage (the MMAs have finished reading it)
            tc::mbar_wait_relaxed(&done_a[g], (it / G) & 1);
            tc::tc_fence_after();
            if (pf) pp[4] = clock64();
#pragma unroll
            for (int mb = 0; mb < 2; ++mb) {
                float z[32];
                uint32_t r0, r1, r2, r3;
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (g * 2 + mb) * NZ;
                tc::tmem_ld32(taddr, z);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr + 32) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int p = mb * 128 + quad * 32 + lane;
                if (p < NPX) {
#pragma unroll
                    for (int tap = 0; tap < 8; ++tap)
                        st4(reinterpret_cast<float*>(stage + (tap * NPX + p) * 16), make_float4(z[tap * 4], z[tap * 4 + 1], z[tap * 4 + 2], z[tap * 4 + 3]));
                    st4(reinterpret_cast<float*>(stage + (8 * NPX + p) * 16),
                        make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), __uint_as_float(r3)));
                }
            }
            tc::tc_fence_before();
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(GROUP_THREADS) : "memory");
            if (pf) pp[5] = clock64();
            float4 o = make_float4(a.bias[0], a.bias[1], a.bias[2], a.bias[3]);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const float4 t = ld4(reinterpret_cast<const float*>(stage + tap * NPX * 16 + q0 + ((tap / 3) * HC + (tap % 3)) * 16));
                o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
            }
            const int y = ty * TR + (lt >> 3), x = tx * TCOL + (lt & 7);
            float s = 0.f, ss = 0.f;
            const int slot = k * a.S + v;                               // (plane, view) slot of t0 / stats
            if (y < a.H && x < a.W) {
                st4(a.t0 + ((size_t)slot * P + (size_t)y * a.W + x) * 4, o);
                s = (o.x + o.y) + (o.z + o.w);
                ss = (o.x * o.x + o.y * o.y) + (o.z * o.z + o.w * o.w);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, off);
                ss += __shfl_xor_sync(0xffffffffu, ss, off);
            }
            if (lane == 0) {       // this warp's private accumulator: a tile contributes 32 pixels x 4 values, ~100 tiles per slot
                float2 acc = sstat[warp - 1][slot];
                acc.x += s; acc.y += ss;
                sstat[warp - 1][slot] = acc;
            }
            if (pf) pp[6] = clock64();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<TMEM_COLS>(tmem_base);
    if (tid < a.K * a.S) {                                                   // stats[k][v][GN1][sum, sumsq]
        double s = 0.0, ss = 0.0;
        for (int w = 0; w < G * 4; ++w) { s += (double)sstat[w][tid].x; ss += (double)sstat[w][tid].y; }
        atomicAdd(a.stats + tid * 6 + 0, s);
        atomicAdd(a.stats + tid * 6 + 1, ss);
    }
}

int g_dbg = 0;
unsigned long long* g_stem_prof = nullptr;

}  // namespace

extern "C" void aarmvs_debug_flags(int flags) { g_dbg = flags; }
extern "C" void aarmvs_debug_stem_profile_buffer(void* buf) { g_stem_prof = (unsigned long long*)buf; }

extern "C" int aarmvs_cost_stem_tc(const float* ref, const float* src, size_t view_stride, int H, int W, int S,
                                   const float* cams, const void* stem_w_f16, const float* stem_b_host,
                                   const float* depth_values, int d_host, const int* d_dev, int planes, int align_corners,
                                   float* t0, double* stats, float* costv, int costv_mode, void* stream) {
    AARMVS_CHECK_ARG(t0 && stats && stem_w_f16 && stem_b_host && depth_values && H > 0 && W > 0);
    AARMVS_CHECK_ARG(costv_mode >= 0 && costv_mode <= 2 && (costv_mode == 0 || costv));
    AARMVS_CHECK_ARG(costv_mode == 2 || (ref && src && cams));
    AARMVS_CHECK_ARG(S >= 1 && S <= AARMVS_MAX_SRC && planes >= 1 && planes <= AARMVS_MAX_PLANES);
    StemArgs a = {};
    a.K = planes;
    a.ref = ref; a.src = src; a.view_stride = view_stride; a.H = H; a.W = W; a.S = S; a.cams = cams;
    a.weight = (const __half*)stem_w_f16;
    for (int i = 0; i < 4; ++i) a.bias[i] = stem_b_host[i];
    a.depth_values = depth_values; a.d_host = d_host; a.d_dev = d_dev; a.align_corners = align_corners; a.t0 = t0;
    a.stats = stats;
    a.costv = costv; a.costv_mode = costv_mode;
    a.dbg = g_dbg;
    a.prof = g_stem_prof;
    static unsigned long long smem_ok = 0;
    if (cudaError_t e = ensure_dyn_smem(cost_stem_tc_kernel, SMEM, smem_ok)) return (int)e;
    const int ntiles = ((W + TCOL - 1) / TCOL) * ((H + TR - 1) / TR);
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    cost_stem_tc_kernel<<<ntiles < sms ? ntiles : sms, N_THREADS, SMEM, (cudaStream_t)stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}
