// ConvLSTM cell as a TMA-fed tcgen05 / TMEM implicit GEMM (models/module.py:76-92).
//
//   gates[128 px][4*CH] = sum over 9 taps, K = Cin:  A_tap[128 px][Cin] * W_tap[Cin][4*CH]
//
// Precision: the parity bar (same winning plane on >= 99.9 % of the pixels) rules out single-pass TF32/BF16 operands
// (hard part 1), so every fp32 operand x travels as two fp16 values x ~= hi + lo
// (|x - hi - lo| <= 2^-22 |x|) and each K=16 step issues kind::f16 MMAs with fp32 accumulation in TMEM:
//   A_hi x [W_hi | W_lo]  (one instruction, N = 2*4CH)   +   A_lo x W_hi        (the dropped lo*lo term is <= 2^-22)
// Useful FLOPs are counted once; the tensor pipe executes 3x as many.
//
// Operand planes.  Every tensor a cell reads lives in HBM ALREADY SPLIT, as "operand planes"
//   [2 (hi, lo)][C/8 chunks][H][W][8 x fp16]
// written by its producer (pass D for the cost, the cell epilogues for h and max-pool(h), aarmvs_split_planes for the
// GroupNorm+ReLU'd deconvolution outputs).  One 16-byte element = 8 channels of one pixel = one row of a K-major
// core matrix, so a TMA box of (10 pixels x 18 rows x chunks) lands in shared memory as exactly the un-swizzled
// K-major operand [chunk][18 x 10 halo pixels][16 B] the MMA descriptors expect: no loader warps, no conversion, no
// LSU traffic (the first version's per-pixel fp32 loads and on-the-fly split kept the L1 data pipe 81 % busy and
// the tensor pipe 45 % busy).  TMA's out-of-bounds zero fill is the conv's zero padding.
//
// Tile: 16 rows x 8 columns of pixels = the 128 rows of the accumulator.  Row m = 8 g + i of the MMA is pixel
// (row g, column i), so one 8-row core-matrix group is one tile row, and inside the staged halo tile
// [18 rows][10 columns] the stride between groups is 10 pixels: a 3x3 tap (ky, kx) is just a different START ADDRESS
// (ky * 10 + kx pixels further) into the same staged data -- no im2col copy, each input pixel is staged once per tile.
// SBO = 10 * 16 B, LBO = 180 * 16 B (validated by aarmvs_tc_selftest).
//
// Warp roles (persistent CTA, one per SM): warps 0-15 epilogue (TMEM -> gates -> c, h; four warps per TMEM lane quadrant,
// each a quarter of the gate columns), warp 16 TMEM allocation + the single MMA-issuing thread, warp 17 the TMA producer.  NST operand stages (producer <-> MMA) and 2 TMEM accumulators
// (MMA <-> epilogue), all hand-offs through mbarriers; tcgen05.commit releases operand stages and publishes
// accumulators.  The cell state c is private to this kernel and kept in a tile-blocked layout
// [tile][quarter of the hidden channels][128 px][CH/4] so that every access is a fully used 128-byte line.
#include "common.cuh"
#include "tc_common.cuh"
#include "plane_map.cuh"
#include "../../include/aarmvs.h"

namespace {

constexpr int TR = 16, TCOL = 8, HR = TR + 2, HC = TCOL + 2, NPX = HR * HC;   // 180 staged pixels per tile
constexpr int CHUNK_BYTES = NPX * 16;                                         // one 8-channel chunk of one fp16 plane
static_assert(HR == planemap::BOX_ROWS && HC == planemap::BOX_COLS, "the tensor-map box is the staged halo tile");
constexpr int N_EPI_WARPS = 16, MMA_WARP = N_EPI_WARPS, TMA_WARP = MMA_WARP + 1;
constexpr int N_THREADS = (N_EPI_WARPS + 2) * 32;                             // 576

struct alignas(64) CellArgs {
    CUtensorMap map[3];         // per source: (W*8 halfs, H, chunks, 2 = hi/lo, ping-pong halves), box (80, 18, chunks, 1, 1)
    int nsrc;                   // sources incl. h_{d-1} (always the last one)
    int chunks[3];              // 16-byte K chunks each source contributes
    int pp_phase[3];            // < 0: single buffer; else the half to read is (d + pp_phase) & 1
    int need_zero;              // K is padded beyond the sources' channels: the pad chunks of every stage must be zero
    int H, W;
    const __half* weight;       // [K steps][2 chunks][2 (hi, lo)][NG][8] fp16 (Cfg::NSTEPS steps)
    const float* bias;          // [NG], order [hidden channel][i,f,o,g]
    __half* h_planes;           // [2 ping-pong][2][CH/8][H][W][8]: plane d writes half d & 1
    float* c_blocked;           // [tiles][4 quarters][128][CH/4]
    float* h_f32;               // optional [2 ping-pong][H*W][CH] fp32 copy of h (deconvolution / head inputs, operator seam)
    __half* pooled;             // optional [2][CH/8][H/2][W/2][8]: operand planes of max_pool2d(h, 2)
    float* const* gates_out;    // optional (training forward): device slot holding the base of the history [planes][H*W][4*CH] of the
                                // ACTIVATED gates, blocks i | f | o | g
    int d_host;
    const int* d_dev;
#ifdef AARMVS_DEBUG
    unsigned long long* prof;   // debug: per-tile clock64 stamps of CTA 0 ([tile][12]), or null
    int dbg;                    // debug ablation: 1 no TMA loads, 4 epilogue skips global I/O, 8 skip MMAs, 16 no transcendentals, 32 no pooling
#endif
};

// Debug hooks (scripts/cell_timeline.py) exist only in -DAARMVS_DEBUG builds: the product library carries neither the
// process-global switches nor their branches in the hot loops.
#ifdef AARMVS_DEBUG
#define CELL_DBG(a, flag) (((a).dbg & (flag)) != 0)
#define CELL_PROF(a) ((a).prof)
#else
#define CELL_DBG(a, flag) (0)
#define CELL_PROF(a) ((unsigned long long*)nullptr)
#endif

template <int KSTEPS, int NG, int NST>
struct Cfg {
    static constexpr int KC8 = 2 * KSTEPS;                       // 16-byte K chunks (8 channels each)
    // Cell 4 (40 input channels = 5 chunks): instead of padding every tap to 3 K steps with a zero chunk, the fifth chunk
    // of TWO taps shares one K = 16 step (the second chunk of an MMA is just a leading-byte-offset away: the next tap's
    // pixel): 9 x 2 + 5 = 23 K steps instead of 27.
    static constexpr bool PACK = (KSTEPS == 3 && NG == 32);
    static constexpr int NSTEPS = PACK ? 23 : 9 * KSTEPS;
    static constexpr int A_PLANE = KC8 * CHUNK_BYTES;            // bytes of one fp16 plane (hi or lo) of one stage
    static constexpr int A_STAGE = 2 * A_PLANE;
    static constexpr int B_BYTES = NSTEPS * 2 * 2 * NG * 16;     // [K step][2 chunks][hi rows | lo rows][16 B]
    static constexpr int SMEM = B_BYTES + NST * A_STAGE + 1024;
    static constexpr int TMEM_COLS = 4 * NG;                     // two accumulators x (hi*hi + lo*hi | hi*lo) column blocks
    static_assert(TMEM_COLS == 128 || TMEM_COLS == 256, "TMEM allocation must be a power of two >= 32");
    static_assert(B_BYTES % 128 == 0 && A_PLANE % 128 == 0, "TMA destinations must be 128-byte aligned");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
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

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// e^{-x} and e^{-2x} in the exp2 domain, clamped so that the product of four (1 + e) terms stays finite (2^120):
// sigmoid(x) >= 2^-30 instead of underflowing further, an absolute error below 1e-9.
__device__ __forceinline__ float exp_neg(float x) { return tc::ex2_approx(fminf(-1.4426950408889634f * x, 30.0f)); }
__device__ __forceinline__ float exp_neg2(float x) { return tc::ex2_approx(fminf(-2.8853900817779268f * x, 30.0f)); }

// SAVE: the training instantiation, which also stores the activated gates (CellArgs::gates_out); the inference one is unchanged.
template <int KSTEPS, int NG, int NST, bool SAVE>
__global__ void __launch_bounds__(N_THREADS, 1) convlstm_tma_kernel(const __grid_constant__ CellArgs a) {
    using C = Cfg<KSTEPS, NG, NST>;
    constexpr int CH = NG / 4, CHK = (CH + 7) / 8;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;                                   // weights: [tap][kc8][hi rows | lo rows][16 B]
    unsigned char* sA = smem + C::B_BYTES;                      // stages: [stage][hi|lo][kc8][px][16 B]
    __shared__ uint64_t full_a[NST], empty_a[NST], tmem_full[2], tmem_empty[2], w_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_bias[NG];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- setup that reads nothing a predecessor kernel writes: under programmatic dependent launch it runs while the
    //      previous kernel of the plane chain drains (barriers, TMEM, tensor-map prefetch, the weight bulk load)
    if (tid < NG) s_bias[tid] = a.bias[tid];
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { tc::mbar_init(&full_a[s], 1); tc::mbar_init(&empty_a[s], 1); }
        for (int b = 0; b < 2; ++b) { tc::mbar_init(&tmem_full[b], 1); tc::mbar_init(&tmem_empty[b], N_EPI_WARPS * 32); }
        tc::mbar_init(&w_bar, 1);
        tc::mbar_fence_init();
    }
    if (warp == MMA_WARP) tc::tmem_alloc<C::TMEM_COLS>(&tmem_base_s);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (warp == TMA_WARP && lane == 0) {
        for (int s = 0; s < a.nsrc; ++s) tc::tma_prefetch_desc(&a.map[s]);
        tc::mbar_arrive_expect_tx(&w_bar, C::B_BYTES);
        for (int off = 0; off < C::B_BYTES; off += 32768)
            tc::bulk_load_1d(sB + off, reinterpret_cast<const unsigned char*>(a.weight) + off, min(32768, C::B_BYTES - off), &w_bar);
    }
    tc::pdl_wait();                          // the predecessor (and, transitively, everything before it) is complete from here on
    tc::pdl_launch_dependents();
    const int d = plane_index(a.d_host, a.d_dev);
    // zeroed stages where no TMA will ever write (h_{-1} = 0 at plane 0, module.py:94-96; K padding of cell 4)
    if (d == 0 || a.need_zero) {
        for (int i = tid; i < NST * C::A_STAGE / 16; i += N_THREADS) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0u, 0u, 0u, 0u);
        tc::fence_async_smem();              // the zero fill went through the generic proxy
        __syncthreads();
    }
    const uint32_t tmem_base = tmem_base_s;

    const int tiles_x = (a.W + TCOL - 1) / TCOL, tiles_y = (a.H + TR - 1) / TR;
    const int ntiles = tiles_x * tiles_y;

    if (warp == TMA_WARP) {
        // =========================================== TMA producer ===========================================
        int cbase[3], sel[3];
        const int nload = d == 0 ? a.nsrc - 1 : a.nsrc;            // plane 0: h_{-1} = 0, its chunks stay zero
        uint32_t bytes = 0;
        {
            int cb = 0;
            for (int s = 0; s < 3; ++s) {
                cbase[s] = cb;
                sel[s] = 0;
                if (s < a.nsrc) {
                    sel[s] = a.pp_phase[s] < 0 ? 0 : ((d + a.pp_phase[s]) & 1);
                    if (s < nload) bytes += 2u * CHUNK_BYTES * a.chunks[s];
                    cb += a.chunks[s];
                }
            }
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int st = it % NST;
            const bool pf = CELL_PROF(a) && blockIdx.x == 0 && lane == 0 && it < 64;
            if (lane == 0) {
                if (pf) CELL_PROF(a)[it * 12 + 0] = clock64();
                tc::mbar_wait(&empty_a[st], ((it / NST) & 1) ^ 1);
                if (pf) CELL_PROF(a)[it * 12 + 1] = clock64();
                if CELL_DBG(a, 1) tc::mbar_arrive(&full_a[st]);
                else tc::mbar_arrive_expect_tx(&full_a[st], bytes);
            }
            __syncwarp();
            // one lane per (source, hi / lo plane): the up to six tensor copies of a tile are issued by ONE warp instruction
            // (a single thread issuing them back to back spent ~400 cycles per copy: 2 400 cycles per tile for cell 4,
            // as long as its MMAs)
            if (lane < 2 * nload && !CELL_DBG(a, 1)) {
                const int s = lane >> 1, hl = lane & 1;
                const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
                unsigned char* dst = sA + st * C::A_STAGE + hl * C::A_PLANE + cbase[s] * CHUNK_BYTES;
                tc::tma_load_5d(dst, &a.map[s], &full_a[st], (tx * TCOL - 1) * 8, ty * TR - 1, 0, hl, sel[s]);
            }
            if (pf) CELL_PROF(a)[it * 12 + 2] = CELL_PROF(a)[it * 12 + 3] = clock64();
            __syncwarp();
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
        tc::mbar_wait(&w_bar, 0);
        // The barriers of tile i+1 are resolved while the MMAs of tile i are still being issued (the issuing thread is
        // throttled by the tensor pipe anyway), so the pipe does not drain between tiles: neither barrier of tile i+1
        // depends on tile i (its stage was released NST tiles earlier, its accumulator two tiles earlier).
        int it = 0;
        if ((int)blockIdx.x < ntiles) {
            tc::mbar_wait(&full_a[0], 0);
            tc::mbar_wait(&tmem_empty[0], 1);
            tc::tc_fence_after();
        }
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int st = it % NST, ab = it & 1;
            const bool pf = CELL_PROF(a) && blockIdx.x == 0 && leader && it < 64;
            if (pf) CELL_PROF(a)[it * 12 + 4] = CELL_PROF(a)[it * 12 + 5] = CELL_PROF(a)[it * 12 + 6] = clock64();
            const uint32_t dst = tmem_base + ab * 2 * NG;
            const uint32_t a_lo_hi = tc::desc_lo(a_base + st * C::A_STAGE, NPX * 16);
            const uint32_t a_lo_lo = tc::desc_lo(a_base + st * C::A_STAGE + C::A_PLANE, NPX * 16);
            const bool more = tile + (int)gridDim.x < ntiles;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                if (leader && !CELL_DBG(a, 8)) {
                    constexpr int KS_TAP = C::PACK ? 2 : KSTEPS;            // full two-chunk K steps of one tap
#pragma unroll
                    for (int ks = 0; ks < KS_TAP; ++ks) {
                        const uint32_t ak = (uint32_t)(((2 * ks) * NPX + (tap / 3) * HC + (tap % 3)) * 16) >> 4;
                        // A_hi x [W_hi | W_lo] fills columns [0, NG) (hi*hi) and [NG, 2NG) (hi*lo) in ONE instruction with
                        // N = 2 NG, A_lo x W_hi adds into [0, NG): two A-operand reads per K step instead of three
                        const uint32_t bk = (uint32_t)((tap * KS_TAP + ks) * 2 * 2 * NG);
                        tc::mma_f16_parts(dst, a_lo_hi + ak, A_HI, b_lo + bk, B_HI, idesc2, (tap | ks) != 0);
                        tc::mma_f16_parts(dst, a_lo_lo + ak, A_HI, b_lo + bk, B_HI, idesc1, 1);
                    }
                    if constexpr (C::PACK) {
                        if (tap % 2 == 0) {        // chunk 4 of taps (tap, tap + 1) in one K step; tap 8 pairs with itself x zero weights
                            constexpr int CH4 = 4 * NPX;
                            const int o0 = (tap / 3) * HC + (tap % 3);
                            const int o1 = tap == 8 ? o0 : ((tap + 1) / 3) * HC + ((tap + 1) % 3);
                            const uint32_t sa = a_base + st * C::A_STAGE + (CH4 + o0) * 16;
                            const uint32_t bk = (uint32_t)((18 + tap / 2) * 2 * 2 * NG);
                            tc::mma_f16_parts(dst, tc::desc_lo(sa, (o1 - o0) * 16), A_HI, b_lo + bk, B_HI, idesc2, 1);
                            tc::mma_f16_parts(dst, tc::desc_lo(sa + C::A_PLANE, (o1 - o0) * 16), A_HI, b_lo + bk, B_HI, idesc1, 1);
                        }
                    }
                }
                if (tap == 5 && more) {          // next tile's operands and accumulator
                    tc::mbar_wait(&full_a[(it + 1) % NST], ((it + 1) / NST) & 1);
                    tc::mbar_wait(&tmem_empty[(it + 1) & 1], (((it + 1) >> 1) & 1) ^ 1);
                    tc::tc_fence_after();
                }
            }
            if (leader) {
                tc::mma_commit(&empty_a[st]);          // operand stage free once these MMAs have read it
                tc::mma_commit(&tmem_full[ab]);        // accumulator complete
            }
            if (pf) CELL_PROF(a)[it * 12 + 7] = clock64();
            __syncwarp();
        }
    } else {
        // =========================================== epilogue ===========================================
        const int q = warp & 3, part = warp >> 2;                // TMEM lane quadrant, quarter of the gate columns
        constexpr int NCOL = NG / 4, NHC = CH / 4;               // gate columns / hidden channels per thread: 16 / 4 or 8 / 2
        const int m = q * 32 + lane;
        const size_t P = (size_t)a.H * a.W;
        __half* hb = a.h_planes + (size_t)(d & 1) * 2 * CHK * P * 8;
        float* hf32 = a.h_f32 ? a.h_f32 + (size_t)(d & 1) * P * CH : nullptr;
        // this thread's NHC channels inside the operand planes: chunk (8 channels) and offset inside its 16-byte element
        const int chunk = (part * NHC) >> 3, sub = (part * NHC) & 7;
        float bias_r[NCOL];
#pragma unroll
        for (int j = 0; j < NCOL; ++j) bias_r[j] = s_bias[part * NCOL + j];
        // previous cell state of this thread's pixel of tile `t` (zero at plane 0 and outside the image)
        auto load_c = [&](int t, float (&cv)[NHC]) {
            const int ty = t / tiles_x, tx = t - ty * tiles_x;
            const bool ok = t < ntiles && ty * TR + (m >> 3) < a.H && tx * TCOL + (m & 7) < a.W && d != 0 && !CELL_DBG(a, 4);
            const float* p = a.c_blocked + (((size_t)t * 4 + part) * 128 + m) * NHC;
            if constexpr (NHC == 4) {
                const float4 v = ok ? ld4(p) : make_float4(0.f, 0.f, 0.f, 0.f);
                cv[0] = v.x; cv[1] = v.y; cv[2] = v.z; cv[3] = v.w;
            } else {
                const float2 v = ok ? *reinterpret_cast<const float2*>(p) : make_float2(0.f, 0.f);
                cv[0] = v.x; cv[1] = v.y;
            }
        };
        float cnext[NHC];
        load_c(blockIdx.x, cnext);
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            const bool pf = CELL_PROF(a) && blockIdx.x == 0 && tid == 0 && it < 64;
            const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
            const int y = ty * TR + (m >> 3), x = tx * TCOL + (m & 7);
            const bool valid = y < a.H && x < a.W && !CELL_DBG(a, 4);
            const size_t pix = valid ? (size_t)y * a.W + x : 0;
            float* cp = a.c_blocked + (((size_t)tile * 4 + part) * 128 + m) * NHC;
            float cprev[NHC], cn[NHC], hn[NHC];
#pragma unroll
            for (int j = 0; j < NHC; ++j) cprev[j] = cnext[j];
            // the NEXT tile's state is requested now: its global-load latency hides behind this tile's gate math (with the
            // small-K cells the accumulator of the next tile is ready before this one's epilogue ends)
            load_c(tile + gridDim.x, cnext);
            if (pf) CELL_PROF(a)[it * 12 + 8] = clock64();
            tc::mbar_wait_relaxed(&tmem_full[ab], (it >> 1) & 1);
            tc::tc_fence_after();
            if (pf) CELL_PROF(a)[it * 12 + 9] = clock64();
            float g[NCOL], g2[NCOL];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 2 * NG + part * NCOL;
            if constexpr (NCOL == 16) { tmem_ld16(taddr, g); tmem_ld16(taddr + NG, g2); }
            else { tmem_ld8(taddr, g); tmem_ld8(taddr + NG, g2); }
            tc::tc_fence_before();
            tc::mbar_arrive(&tmem_empty[ab]);                    // accumulator is in registers: release it
            if (pf) CELL_PROF(a)[it * 12 + 10] = clock64();
            // pixels outside the image (ragged tiles) see zero operands: finite values that are never stored.
            // Gates with ONE reciprocal per hidden channel: with a = 1 + e^-i, b = 1 + e^-f, c = 1 + e^-o, d = 1 + e^-2g and
            // r = 1 / (abcd):  sigmoid(i) = r*cd*b, sigmoid(f) = r*cd*a, sigmoid(o) = r*ab*d, tanh(g) = 2*r*ab*c - 1
            // (MUFU.EX2 / MUFU.RCP are quarter-rate: 6.5 instead of 10 per channel; relative error a few ulp).
            float ec[NHC], go[NHC], gia[NHC], gfa[NHC], gga[NHC];
#pragma unroll
            for (int j = 0; j < NHC; ++j) {
                const float* bb = bias_r + j * 4;
                if CELL_DBG(a, 16) {          // debug ablation: no transcendental math
                    cn[j] = (g[j * 4 + 0] + g2[j * 4 + 0]) + bb[0] + cprev[j];
                    go[j] = (g[j * 4 + 1] + g2[j * 4 + 1]) + (g[j * 4 + 2] + g2[j * 4 + 2]) + (g[j * 4 + 3] + g2[j * 4 + 3]);
                    ec[j] = 1.0f;
                    if constexpr (SAVE) gia[j] = gfa[j] = gga[j] = 0.f;
                    continue;
                }
                const float pa = 1.0f + exp_neg((g[j * 4 + 0] + g2[j * 4 + 0]) + bb[0]);      // (hi*hi + lo*hi) + hi*lo, + bias
                const float pb = 1.0f + exp_neg((g[j * 4 + 1] + g2[j * 4 + 1]) + bb[1]);
                const float pc = 1.0f + exp_neg((g[j * 4 + 2] + g2[j * 4 + 2]) + bb[2]);
                const float pd = 1.0f + exp_neg2((g[j * 4 + 3] + g2[j * 4 + 3]) + bb[3]);
                const float pab = pa * pb, pcd = pc * pd;
                const float r = tc::rcp_approx(pab * pcd);
                const float rab = r * pcd, rcd = r * pab;                                    // 1/(ab), 1/(cd)
                const float gi = rab * pb, gf = rab * pa;
                go[j] = rcd * pd;
                const float gg = fmaf(2.0f, rcd * pc, -1.0f);
                cn[j] = __fadd_rn(__fmul_rn(gf, cprev[j]), __fmul_rn(gi, gg));
                ec[j] = 1.0f + exp_neg2(cn[j]);
                if constexpr (SAVE) { gia[j] = gi; gfa[j] = gf; gga[j] = gg; }
            }
#pragma unroll
            for (int j = 0; j < NHC; j += 2) {           // tanh(c') of two channels with one reciprocal
                const float r = tc::rcp_approx(ec[j] * ec[j + 1]);
                hn[j] = go[j] * fmaf(2.0f, r * ec[j + 1], -1.0f);
                hn[j + 1] = go[j + 1] * fmaf(2.0f, r * ec[j], -1.0f);
            }
            if (SAVE && valid) {             // the gates exactly as this forward used them, for the backward pass (no recompute there)
                float* gp = *a.gates_out + ((size_t)d * P + pix) * (4 * CH) + part * NHC;
                if constexpr (NHC == 4) {
                    st4(gp, make_float4(gia[0], gia[1], gia[2], gia[3]));
                    st4(gp + CH, make_float4(gfa[0], gfa[1], gfa[2], gfa[3]));
                    st4(gp + 2 * CH, make_float4(go[0], go[1], go[2], go[3]));
                    st4(gp + 3 * CH, make_float4(gga[0], gga[1], gga[2], gga[3]));
                } else {
                    *reinterpret_cast<float2*>(gp) = make_float2(gia[0], gia[1]);
                    *reinterpret_cast<float2*>(gp + CH) = make_float2(gfa[0], gfa[1]);
                    *reinterpret_cast<float2*>(gp + 2 * CH) = make_float2(go[0], go[1]);
                    *reinterpret_cast<float2*>(gp + 3 * CH) = make_float2(gga[0], gga[1]);
                }
            }
            if (valid) {
                if constexpr (NHC == 4) {
                    st4(cp, make_float4(cn[0], cn[1], cn[2], cn[3]));
                    uint2 h0, l0;
                    tc::split4_f16(make_float4(hn[0], hn[1], hn[2], hn[3]), h0, l0);
                    *reinterpret_cast<uint2*>(hb + ((size_t)chunk * P + pix) * 8 + sub) = h0;
                    *reinterpret_cast<uint2*>(hb + ((size_t)(CHK + chunk) * P + pix) * 8 + sub) = l0;
                    if (hf32) st4(hf32 + pix * CH + part * NHC, make_float4(hn[0], hn[1], hn[2], hn[3]));
                } else {
                    *reinterpret_cast<float2*>(cp) = make_float2(cn[0], cn[1]);
                    __half hh0, hl0, hh1, hl1;
                    tc::split_f16_trunc(hn[0], hh0, hl0);
                    tc::split_f16_trunc(hn[1], hh1, hl1);
                    *reinterpret_cast<__half2*>(hb + ((size_t)chunk * P + pix) * 8 + sub) = __halves2half2(hh0, hh1);
                    *reinterpret_cast<__half2*>(hb + ((size_t)(CHK + chunk) * P + pix) * 8 + sub) = __halves2half2(hl0, hl1);
                    if (hf32) *reinterpret_cast<float2*>(hf32 + pix * CH + part * NHC) = make_float2(hn[0], hn[1]);
                }
            }
            if constexpr (NHC == 4) {
                if (a.pooled && !CELL_DBG(a, 32)) {    // max_pool2d(h, 2) (drmvsnet.py:148,152): the 2x2 block is lanes ^1, ^8 of this warp
                    float pm[NHC];
#pragma unroll
                    for (int j = 0; j < NHC; ++j) {
                        float v = fmaxf(hn[j], __shfl_xor_sync(0xffffffffu, hn[j], 1));
                        pm[j] = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
                    }
                    if (valid && !(lane & 9)) {
                        const size_t P4 = P >> 2, pp = (size_t)(y >> 1) * (a.W >> 1) + (x >> 1);
                        uint2 h0, l0;
                        tc::split4_f16(make_float4(pm[0], pm[1], pm[2], pm[3]), h0, l0);
                        *reinterpret_cast<uint2*>(a.pooled + ((size_t)chunk * P4 + pp) * 8 + sub) = h0;
                        *reinterpret_cast<uint2*>(a.pooled + ((size_t)(CHK + chunk) * P4 + pp) * 8 + sub) = l0;
                    }
                }
            }
            if (pf) CELL_PROF(a)[it * 12 + 11] = clock64();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int KSTEPS, int NG, int NST, bool SAVE>
int launch_cell_as(const CellArgs& a, cudaStream_t stream) {
    using C = Cfg<KSTEPS, NG, NST>;
    auto kern = convlstm_tma_kernel<KSTEPS, NG, NST, SAVE>;
    static unsigned long long smem_ok = 0;
    if (cudaError_t e = ensure_dyn_smem(kern, C::SMEM, smem_ok)) return (int)e;
    const int ntiles = ((a.W + TCOL - 1) / TCOL) * ((a.H + TR - 1) / TR);
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    if (cudaError_t e = launch_pdl(kern, dim3(balanced_grid(ntiles, sms)), dim3(N_THREADS), C::SMEM, stream, a)) return (int)e;
    AARMVS_LAUNCH_RESULT();
}

template <int KSTEPS, int NG, int NST>
int launch_cell(const CellArgs& a, cudaStream_t stream) {
    return a.gates_out ? launch_cell_as<KSTEPS, NG, NST, true>(a, stream) : launch_cell_as<KSTEPS, NG, NST, false>(a, stream);
}

#ifdef AARMVS_DEBUG
unsigned long long* g_prof = nullptr;
int g_cell_dbg = 0;
#endif

}  // namespace

#ifdef AARMVS_DEBUG
// Debug only: CTA 0 of the next ConvLSTM launches writes clock64 stamps per tile ([64][12] uint64) into `buf` (device).
extern "C" void aarmvs_debug_profile_buffer(void* buf) { g_prof = (unsigned long long*)buf; }
extern "C" void aarmvs_debug_cell_flags(int flags) { g_cell_dbg = flags; }
#endif

extern "C" size_t aarmvs_cell_state_floats(int H, int W, int hidden) {
    if (H <= 0 || W <= 0 || hidden <= 0) return 0;
    return (size_t)((W + TCOL - 1) / TCOL) * ((H + TR - 1) / TR) * 128 * hidden;
}

extern "C" int aarmvs_convlstm_cell_tc(const aarmvs_plane_src* src, int nsrc, int H, int W, int cin, int hidden,
                                       const void* weight_f16, const float* bias_packed, void* h_planes, float* c_blocked,
                                       float* h_f32_pair, void* pooled_planes, float* const* gates_out, int d_host, const int* d_dev,
                                       void* stream) {
    AARMVS_CHECK_ARG(src && weight_f16 && bias_packed && h_planes && c_blocked && H > 0 && W > 0 && nsrc >= 1 && nsrc <= 2);
    AARMVS_CHECK_ARG(hidden == 16 || hidden == 8);
    AARMVS_CHECK_ARG(!pooled_planes || (hidden == 16 && H % 2 == 0 && W % 2 == 0));
    CellArgs a = {};
    a.nsrc = nsrc + 1;
    int csum = 0;
    for (int s = 0; s <= nsrc; ++s) {
        const void* planes;
        int ch, pp;
        if (s < nsrc) {
            AARMVS_CHECK_ARG(src[s].planes && src[s].channels > 0 && src[s].channels % 8 == 0 && (src[s].pp == 1 || src[s].pp == 2));
            AARMVS_CHECK_ARG(((uintptr_t)src[s].planes & 15) == 0);
            planes = src[s].planes; ch = src[s].channels; pp = src[s].pp;
            a.pp_phase[s] = pp == 2 ? (src[s].pp_phase & 1) : -1;
        } else {      // h_{d-1}: the half of the ping-pong pair that plane d does not write
            planes = h_planes; ch = hidden; pp = 2;
            a.pp_phase[s] = 1;
        }
        a.chunks[s] = ch / 8;
        if (int e = planemap::make_plane_map(&a.map[s], planes, H, W, ch / 8, pp)) return e;
        csum += ch;
    }
    AARMVS_CHECK_ARG(csum == cin && cin <= 48 && ((uintptr_t)h_planes & 15) == 0 && ((uintptr_t)weight_f16 & 15) == 0);
    a.need_zero = 0;          // K padding is never read: cell 4 packs its fifth chunk (see Cfg::PACK)
    a.H = H; a.W = W; a.weight = (const __half*)weight_f16; a.bias = bias_packed; a.h_planes = (__half*)h_planes;
    a.c_blocked = c_blocked; a.h_f32 = h_f32_pair; a.pooled = (__half*)pooled_planes; a.gates_out = gates_out;
    a.d_host = d_host; a.d_dev = d_dev;
#ifdef AARMVS_DEBUG
    a.prof = g_prof; a.dbg = g_cell_dbg;
#endif
    cudaStream_t st = (cudaStream_t)stream;
    if (cin == 48 && hidden == 16) return launch_cell<3, 64, 3>(a, st);     // cells 0, 3
    if (cin == 32 && hidden == 16) return launch_cell<2, 64, 4>(a, st);     // cells 1, 2
    if (cin == 40 && hidden == 8) return launch_cell<3, 32, 4>(a, st);      // cell 4 (K padded 40 -> 48 with a zero chunk)
    return AARMVS_ERR_ARG;
}
