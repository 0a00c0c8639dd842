// ConvLSTM cell as a tcgen05 / TMEM implicit GEMM (models/module.py:76-92).
//
//   gates[128 px][4*CH] = sum over 9 taps, K = Cin:  A_tap[128 px][Cin] * W_tap[Cin][4*CH]
//
// Precision: the parity bar (same winning plane on >= 99.9 % of the pixels) rules out single-pass TF32/BF16 operands
// (hard part 1), so every fp32 operand x is split on the fly into two fp16 values x ~= hi + lo
// (|x - hi - lo| <= 2^-22 |x|) and each K=16 step issues three kind::f16 MMAs with fp32 accumulation in TMEM:
//   hi*hi + lo*hi + hi*lo      (the dropped lo*lo term is <= 2^-22 relative)
// Useful FLOPs are counted once; the tensor pipe executes 3x as many.
//
// Tile: 16 rows x 8 columns of pixels = the 128 rows of the accumulator.  Row m = 8 g + i of the MMA is pixel
// (row g, column i), so one 8-row core-matrix group is one tile row, and inside the staged halo tile
// [18 rows][10 columns] the stride between groups is 10 pixels: a 3x3 tap (ky, kx) is just a different START ADDRESS
// (ky * 10 + kx pixels further) into the same staged data -- no im2col copy, each input pixel is staged once per tile.
// Shared-memory operand layout (un-swizzled, K-major): [16-byte K chunk = 8 channels][pixel][8 x fp16], i.e.
// 8 pixels x 16 B contiguous per core matrix, SBO = 10 * 16 B, LBO = 180 * 16 B (validated by aarmvs_tc_selftest).
//
// Warp roles (persistent CTA, one per SM):  warps 0-7 epilogue (TMEM -> gates -> c, h),  warp 8 TMEM allocation + the
// single MMA-issuing thread,  warps 9-16 loaders (global fp32 -> concat / GroupNorm+ReLU / max-pool on load -> fp16
// hi/lo split -> shared memory).  Pipelines: 2 operand stages (loader <-> MMA) and 2 TMEM accumulators (MMA <->
// epilogue), all hand-offs through mbarriers; tcgen05.commit releases operand stages and publishes accumulators.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/aarmvs.h"

namespace {

struct TcSrc {
    const float* ptr;
    const double* stats;
    const float* gamma;
    const float* beta;
    int ch, mode, groups;
    double count;
    size_t pp_stride;
    int pp_phase;
};

struct TcArgs {
    TcSrc src[3];
    int nsrc;
    int H, W;
    const __half* weight;      // [2 planes hi,lo][9 taps][KC8][NG][8] fp16
    const float* bias;         // [NG], order [hidden channel][i,f,o,g]
    float* h_pair;
    float* c_state;
    int d_host;
    const int* d_dev;
    unsigned long long* prof;   // debug: per-tile clock64 stamps of CTA 0 ([tile][12]), or null
    int dbg;                    // debug ablation: 1 loader skips global loads, 2 loader skips smem stores, 4 epilogue skips global I/O, 8 skip MMAs
};

constexpr int TR = 16, TCOL = 8, HR = TR + 2, HC = TCOL + 2, NPX = HR * HC;   // 180 staged pixels per tile
constexpr int N_EPI_WARPS = 8, N_LOAD_WARPS = 8;
constexpr int MMA_WARP = N_EPI_WARPS, FIRST_LOAD_WARP = MMA_WARP + 1;
constexpr int N_THREADS = (N_EPI_WARPS + 1 + N_LOAD_WARPS) * 32;               // 544
constexpr int N_STAGES = 2;

template <int KSTEPS, int NG>
struct TcCfg {
    static constexpr int KC8 = 2 * KSTEPS;                       // 16-byte K chunks (8 channels each)
    static constexpr int A_PLANE = KC8 * NPX * 16;               // bytes of one fp16 plane (hi or lo) of one stage
    static constexpr int A_STAGE = 2 * A_PLANE;
    static constexpr int B_BYTES = 9 * KC8 * 2 * NG * 16;        // [tap][chunk][hi rows | lo rows][16 B]
    static constexpr int SMEM = B_BYTES + N_STAGES * A_STAGE + 1024;
    static constexpr int TMEM_COLS = 4 * NG;                     // two accumulators x (hi*hi + lo*hi | hi*lo) column blocks
    static_assert(TMEM_COLS == 128 || TMEM_COLS == 256, "TMEM allocation must be a power of two >= 32");
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

template <int KSTEPS, int NG>
__global__ void __launch_bounds__(N_THREADS, 1) convlstm_tc_kernel(const __grid_constant__ TcArgs a) {
    using C = TcCfg<KSTEPS, NG>;
    constexpr int CH = NG / 4;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;                                   // weights: [hi|lo][tap][kc8][n][16 B]
    unsigned char* sA = smem + C::B_BYTES;                      // stages: [stage][hi|lo][kc8][px][16 B]
    __shared__ uint64_t full_a[N_STAGES], empty_a[N_STAGES], tmem_full[2], tmem_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_scale[48], s_shift[48], s_bias[NG];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int d = plane_index(a.d_host, a.d_dev);

    // ---- one-time setup: weights -> smem, GroupNorm scale/shift of normalised sources, barriers, TMEM
    for (int i = tid; i < C::B_BYTES / 16; i += N_THREADS)
        reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(a.weight) + i);
    if (tid < 48) {
        float sc = 1.f, sh = 0.f;
        int base = 0;
        for (int s = 0; s < a.nsrc; ++s) {
            const TcSrc& S = a.src[s];
            if (tid >= base && tid < base + S.ch && S.mode == AARMVS_SRC_GN_RELU) {
                const int cl = tid - base;
                float mean, rstd;
                gn_mean_rstd(S.stats + (cl / (S.ch / S.groups)) * 2, S.count, mean, rstd);
                sc = rstd * S.gamma[cl];
                sh = S.beta[cl] - mean * sc;
            }
            base += S.ch;
        }
        s_scale[tid] = sc;
        s_shift[tid] = sh;
    }
    if (tid < NG) s_bias[tid] = a.bias[tid];
    if (tid == 0) {
        for (int s = 0; s < N_STAGES; ++s) { tc::mbar_init(&full_a[s], N_LOAD_WARPS * 32); tc::mbar_init(&empty_a[s], 1); }
        for (int b = 0; b < 2; ++b) { tc::mbar_init(&tmem_full[b], 1); tc::mbar_init(&tmem_empty[b], N_EPI_WARPS * 32); }
        tc::mbar_fence_init();
    }
    if (warp == MMA_WARP) tc::tmem_alloc<C::TMEM_COLS>(&tmem_base_s);
    tc::fence_async_smem();                  // weights were written through the generic proxy
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    const int tiles_x = (a.W + TCOL - 1) / TCOL, tiles_y = (a.H + TR - 1) / TR;
    const int ntiles = tiles_x * tiles_y;

    if (warp >= FIRST_LOAD_WARP) {
        // =========================================== loaders ===========================================
        // One item = 8 channels (32 B of fp32 -> one 16-byte fp16 chunk per plane) of one staged pixel.  The decoding of a
        // thread's items (source, pixel, channel) does not depend on the tile and is done once; per tile an item costs
        // an address, two (max-pool: eight) 16-byte loads, the split and two conflict-free 16-byte stores.
        const int lt = tid - FIRST_LOAD_WARP * 32;
        constexpr int LT = N_LOAD_WARPS * 32;
        constexpr int ITEMS = (C::KC8 * NPX + LT - 1) / LT;
        const float* sptr[3];
        int sbase[4] = {0, 0, 0, 0};
        for (int s = 0; s < 3; ++s) {
            sptr[s] = nullptr;
            if (s < a.nsrc) {
                const TcSrc& S = a.src[s];
                sptr[s] = S.pp_stride ? S.ptr + (size_t)((d + S.pp_phase) & 1) * S.pp_stride : S.ptr;
                sbase[s + 1] = sbase[s] + S.ch;
            } else {
                sbase[s + 1] = sbase[s];
            }
        }
        // per item: meta = source (2 bits; 3 = zero fill) | row << 2 | col << 7 | local channel << 11 | mode << 17, and
        // the source's channel count
        int meta[ITEMS], sch[ITEMS];
        const float* base[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int idx = lt + k * LT;
            int m = 3, ch = 0;
            base[k] = nullptr;
            if (idx < C::KC8 * NPX) {
                const int kc = idx / NPX, px = idx - kc * NPX;
                const int r = px / HC, cc = px - r * HC, c = kc * 8;
                if (c < sbase[a.nsrc]) {
                    const int s = c < sbase[1] ? 0 : (c < sbase[2] ? 1 : 2);
                    const bool zero_h = (s == a.nsrc - 1 && d == 0);              // h_{-1} = 0 (module.py:94-96)
                    if (!zero_h) {
                        m = s | (r << 2) | (cc << 7) | ((c - sbase[s]) << 11) | (a.src[s].mode << 17);
                        ch = a.src[s].ch;
                        base[k] = (s == 0 ? sptr[0] : s == 1 ? sptr[1] : sptr[2]) + (c - sbase[s]);
                    } else {
                        m = 3 | (r << 2) | (cc << 7);
                    }
                } else {
                    m = 3 | (r << 2) | (cc << 7);
                }
            } else {
                m = -1;                                                             // no item
            }
            meta[k] = m; sch[k] = ch;
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int st = it % N_STAGES;
            const bool pf = a.prof && blockIdx.x == 0 && lt == 0 && it < 64;
            if (pf) a.prof[it * 12 + 0] = clock64();
            const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
            const int y0 = ty * TR - 1, x0 = tx * TCOL - 1;
            // loads first (registers), then wait for the stage, then convert + store
            float4 va[ITEMS], vb[ITEMS];
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                va[k] = vb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int m = meta[k];
                if (m < 0 || (m & 3) == 3) continue;
                const int y = y0 + ((m >> 2) & 31), x = x0 + ((m >> 7) & 15);
                if (y < 0 || y >= a.H || x < 0 || x >= a.W || (a.dbg & 1)) continue;
                const int mode = m >> 17;
                if (mode == AARMVS_SRC_MAXPOOL2) {
                    const size_t rowp = (size_t)2 * a.W * sch[k];
                    const float* p = base[k] + (size_t)(2 * y) * rowp + (size_t)(2 * x) * sch[k];
                    const float4 q0 = ldg4(p), q1 = ldg4(p + sch[k]), q2 = ldg4(p + rowp), q3 = ldg4(p + rowp + sch[k]);
                    const float4 r0 = ldg4(p + 4), r1 = ldg4(p + sch[k] + 4), r2 = ldg4(p + rowp + 4), r3 = ldg4(p + rowp + sch[k] + 4);
                    va[k] = make_float4(fmaxf(fmaxf(q0.x, q1.x), fmaxf(q2.x, q3.x)), fmaxf(fmaxf(q0.y, q1.y), fmaxf(q2.y, q3.y)),
                                        fmaxf(fmaxf(q0.z, q1.z), fmaxf(q2.z, q3.z)), fmaxf(fmaxf(q0.w, q1.w), fmaxf(q2.w, q3.w)));
                    vb[k] = make_float4(fmaxf(fmaxf(r0.x, r1.x), fmaxf(r2.x, r3.x)), fmaxf(fmaxf(r0.y, r1.y), fmaxf(r2.y, r3.y)),
                                        fmaxf(fmaxf(r0.z, r1.z), fmaxf(r2.z, r3.z)), fmaxf(fmaxf(r0.w, r1.w), fmaxf(r2.w, r3.w)));
                } else {
                    const float* p = base[k] + ((size_t)y * a.W + x) * sch[k];
                    va[k] = ldg4(p);
                    vb[k] = ldg4(p + 4);
                    if (mode == AARMVS_SRC_GN_RELU) {          // the conv pads with zeros AFTER GroupNorm + ReLU
                        const int c = sbase[m & 3] + ((m >> 11) & 63);
                        const float* sc = s_scale + c;
                        const float* sh = s_shift + c;
                        va[k].x = fmaxf(fmaf(va[k].x, sc[0], sh[0]), 0.f); va[k].y = fmaxf(fmaf(va[k].y, sc[1], sh[1]), 0.f);
                        va[k].z = fmaxf(fmaf(va[k].z, sc[2], sh[2]), 0.f); va[k].w = fmaxf(fmaf(va[k].w, sc[3], sh[3]), 0.f);
                        vb[k].x = fmaxf(fmaf(vb[k].x, sc[4], sh[4]), 0.f); vb[k].y = fmaxf(fmaf(vb[k].y, sc[5], sh[5]), 0.f);
                        vb[k].z = fmaxf(fmaf(vb[k].z, sc[6], sh[6]), 0.f); vb[k].w = fmaxf(fmaf(vb[k].w, sc[7], sh[7]), 0.f);
                    }
                }
            }
            tc::mbar_wait_relaxed(&empty_a[st], ((it / N_STAGES) & 1) ^ 1);
            if (pf) a.prof[it * 12 + 1] = clock64();
            unsigned char* stage = sA + st * C::A_STAGE;
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                if (meta[k] < 0 || (a.dbg & 2)) continue;
                const int idx = lt + k * LT;
                uint2 h0, l0, h1, l1;
                tc::split4_f16(va[k], h0, l0);
                tc::split4_f16(vb[k], h1, l1);
                *reinterpret_cast<uint4*>(stage + (size_t)idx * 16) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                *reinterpret_cast<uint4*>(stage + C::A_PLANE + (size_t)idx * 16) = make_uint4(l0.x, l0.y, l1.x, l1.y);
            }
            if (pf) a.prof[it * 12 + 2] = clock64();
            tc::fence_async_smem();
            tc::mbar_arrive(&full_a[st]);
            if (pf) a.prof[it * 12 + 3] = clock64();
        }
    } else if (warp == MMA_WARP) {
        // =========================================== MMA issuer ===========================================
        // The whole warp runs the loop (uniform control flow keeps the descriptor arithmetic on the uniform datapath);
        // one elected lane issues.  81 / 54 MMAs per tile, fully unrolled: descriptor low words are base + immediate.
        const bool leader = tc::elect_one();
        const uint32_t idesc2 = tc::idesc_f16_f32(128, 2 * NG), idesc1 = tc::idesc_f16_f32(128, NG);
        const uint32_t a_base = tc::smem_u32(sA), b_base = tc::smem_u32(sB);
        constexpr uint32_t A_HI = tc::desc_hi(HC * 16), B_HI = tc::desc_hi(8 * 16);
        const uint32_t b_lo = tc::desc_lo(b_base, 2 * NG * 16);       // [W_hi | W_lo] stacked along N: rows 0..NG-1 hi, NG..2NG-1 lo
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int st = it % N_STAGES, ab = it & 1;
            const bool pf = a.prof && blockIdx.x == 0 && leader && it < 64;
            if (pf) a.prof[it * 12 + 4] = clock64();
            tc::mbar_wait(&full_a[st], (it / N_STAGES) & 1);
            if (pf) a.prof[it * 12 + 5] = clock64();
            tc::mbar_wait(&tmem_empty[ab], ((it >> 1) & 1) ^ 1);
            tc::tc_fence_after();
            if (pf) a.prof[it * 12 + 6] = clock64();
            const uint32_t dst = tmem_base + ab * 2 * NG;
            const uint32_t a_lo_hi = tc::desc_lo(a_base + st * C::A_STAGE, NPX * 16);
            const uint32_t a_lo_lo = tc::desc_lo(a_base + st * C::A_STAGE + C::A_PLANE, NPX * 16);
            if (leader) {
                if (!(a.dbg & 8))
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks) {
                        const uint32_t ak = (uint32_t)(((2 * ks) * NPX + (tap / 3) * HC + (tap % 3)) * 16) >> 4;
                        // A_hi x [W_hi | W_lo] fills columns [0, NG) (hi*hi) and [NG, 2NG) (hi*lo) in ONE instruction with
                        // N = 2 NG, A_lo x W_hi adds into [0, NG): two A-operand reads per K step instead of three
                        const uint32_t bk = (uint32_t)((tap * C::KC8 + 2 * ks) * 2 * NG);
                        tc::mma_f16_parts(dst, a_lo_hi + ak, A_HI, b_lo + bk, B_HI, idesc2, (tap | ks) != 0);
                        tc::mma_f16_parts(dst, a_lo_lo + ak, A_HI, b_lo + bk, B_HI, idesc1, 1);
                    }
                }
                tc::mma_commit(&empty_a[st]);          // operand stage free once these MMAs have read it
                tc::mma_commit(&tmem_full[ab]);        // accumulator complete
            }
            if (pf) a.prof[it * 12 + 7] = clock64();
            __syncwarp();
        }
    } else {
        // =========================================== epilogue ===========================================
        const int q = warp & 3, hf = warp >> 2;                  // TMEM lane quadrant, half of the gate columns
        constexpr int NCOL = NG / 2, NHC = CH / 2;               // columns / hidden channels per thread
        const int m = q * 32 + lane;
        float* hdst = a.h_pair + (size_t)(d & 1) * a.H * a.W * CH;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            const bool pf = a.prof && blockIdx.x == 0 && tid == 0 && it < 64;
            // this thread's pixel and its previous cell state: requested BEFORE waiting for the accumulator so that the
            // global-load latency overlaps the MMAs of this tile
            const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
            const int y = ty * TR + (m >> 3), x = tx * TCOL + (m & 7);
            const bool valid = y < a.H && x < a.W && !(a.dbg & 4);
            const size_t pix = valid ? (size_t)y * a.W + x : 0;
            float* cp = a.c_state + pix * CH + hf * NHC;
            float* hp = hdst + pix * CH + hf * NHC;
            float cprev[NHC], cn[NHC], hn[NHC];
#pragma unroll
            for (int j = 0; j < NHC; j += 4) {
                const float4 t = d == 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : ld4(cp + j);
                cprev[j] = t.x; cprev[j + 1] = t.y; cprev[j + 2] = t.z; cprev[j + 3] = t.w;
            }
            if (pf) a.prof[it * 12 + 8] = clock64();
            tc::mbar_wait_relaxed(&tmem_full[ab], (it >> 1) & 1);
            tc::tc_fence_after();
            if (pf) a.prof[it * 12 + 9] = clock64();
            float g[NCOL], g2[NCOL];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 2 * NG + hf * NCOL;
            if constexpr (NCOL == 32) { tc::tmem_ld32(taddr, g); tc::tmem_ld32(taddr + NG, g2); }
            else { tmem_ld16(taddr, g); tmem_ld16(taddr + NG, g2); }
#pragma unroll
            for (int j = 0; j < NCOL; ++j) g[j] += g2[j];            // (hi*hi + lo*hi) + hi*lo
            tc::tc_fence_before();
            tc::mbar_arrive(&tmem_empty[ab]);                    // accumulator is in registers: release it
            if (pf) a.prof[it * 12 + 10] = clock64();
            if (valid) {
#pragma unroll
                for (int j = 0; j < NHC; ++j) {
                    const float* bb = s_bias + hf * NCOL + j * 4;
                    const float gi = tc::fast_sigmoid(g[j * 4 + 0] + bb[0]);
                    const float gf = tc::fast_sigmoid(g[j * 4 + 1] + bb[1]);
                    const float go = tc::fast_sigmoid(g[j * 4 + 2] + bb[2]);
                    const float gg = tc::fast_tanh(g[j * 4 + 3] + bb[3]);
                    cn[j] = __fadd_rn(__fmul_rn(gf, cprev[j]), __fmul_rn(gi, gg));
                    hn[j] = go * tc::fast_tanh(cn[j]);
                }
#pragma unroll
                for (int j = 0; j < NHC; j += 4) {
                    st4(cp + j, make_float4(cn[j], cn[j + 1], cn[j + 2], cn[j + 3]));
                    st4(hp + j, make_float4(hn[j], hn[j + 1], hn[j + 2], hn[j + 3]));
                }
            }
            if (pf) a.prof[it * 12 + 11] = clock64();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int KSTEPS, int NG>
int launch_tc(const TcArgs& a, cudaStream_t stream) {
    using C = TcCfg<KSTEPS, NG>;
    auto kern = convlstm_tc_kernel<KSTEPS, NG>;
    static unsigned long long smem_ok = 0;
    if (cudaError_t e = ensure_dyn_smem(kern, C::SMEM, smem_ok)) return (int)e;
    const int ntiles = ((a.W + TCOL - 1) / TCOL) * ((a.H + TR - 1) / TR);
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    kern<<<ntiles < sms ? ntiles : sms, N_THREADS, C::SMEM, stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}

unsigned long long* g_prof = nullptr;
int g_cell_dbg = 0;

}  // namespace

// Debug only: CTA 0 of the next ConvLSTM launches writes clock64 stamps per tile ([64][12] uint64) into `buf` (device).
extern "C" void aarmvs_debug_profile_buffer(void* buf) { g_prof = (unsigned long long*)buf; }
extern "C" void aarmvs_debug_cell_flags(int flags) { g_cell_dbg = flags; }

extern "C" int aarmvs_convlstm_cell_tc(const aarmvs_conv_src* src, int nsrc, int H, int W, int cin, int hidden,
                                       const void* weight_f16, const float* bias_packed, float* h_pair, float* c_state,
                                       int d_host, const int* d_dev, void* stream) {
    AARMVS_CHECK_ARG(src && weight_f16 && bias_packed && h_pair && c_state && H > 0 && W > 0 && nsrc >= 1 && nsrc <= 2);
    TcArgs a = {};
    a.nsrc = nsrc + 1;
    int csum = 0;
    for (int s = 0; s <= nsrc; ++s) {
        TcSrc& t = a.src[s];
        if (s < nsrc) {
            const aarmvs_conv_src& i = src[s];
            AARMVS_CHECK_ARG(i.ptr && i.channels > 0 && i.channels % 8 == 0);
            AARMVS_CHECK_ARG(i.mode == AARMVS_SRC_PLAIN || i.mode == AARMVS_SRC_GN_RELU || i.mode == AARMVS_SRC_MAXPOOL2);
            if (i.mode == AARMVS_SRC_GN_RELU)
                AARMVS_CHECK_ARG(i.gn_sums && i.gn_weight && i.gn_bias && i.gn_groups >= 1 && i.gn_count > 0);
            t.ptr = i.ptr; t.stats = i.gn_sums; t.gamma = i.gn_weight; t.beta = i.gn_bias; t.ch = i.channels; t.mode = i.mode;
            t.groups = i.gn_groups; t.count = i.gn_count; t.pp_stride = i.pp_stride; t.pp_phase = i.pp_phase;
        } else {      // h_{d-1}: the half of the ping-pong pair that plane d does not write
            t.ptr = h_pair; t.ch = hidden; t.mode = AARMVS_SRC_PLAIN; t.pp_stride = (size_t)H * W * hidden; t.pp_phase = 1;
        }
        csum += t.ch;
    }
    AARMVS_CHECK_ARG(csum == cin && cin <= 48);
    a.H = H; a.W = W; a.weight = (const __half*)weight_f16; a.bias = bias_packed; a.h_pair = h_pair; a.c_state = c_state;
    a.d_host = d_host; a.d_dev = d_dev; a.prof = g_prof; a.dbg = g_cell_dbg;
    cudaStream_t st = (cudaStream_t)stream;
    if (cin == 48 && hidden == 16) return launch_tc<3, 64>(a, st);     // cells 0, 3
    if (cin == 32 && hidden == 16) return launch_tc<2, 64>(a, st);     // cells 1, 2
    if (cin == 40 && hidden == 8) return launch_tc<3, 32>(a, st);      // cell 4 (K padded 40 -> 48 with zero channels)
    return AARMVS_ERR_ARG;
}
