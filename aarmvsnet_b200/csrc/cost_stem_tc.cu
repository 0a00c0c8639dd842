// Pass A of the plane sweep on the tensor cores (models/module.py:6-38,
// models/drmvsnet.py:31,311): for every source view, bilinear gather at depth d, c = (warp - ref)^2, and the omega
// stem 3x3 conv 32 -> 4 (omega.reweight_network.0.0) on tcgen05.
//
// The conv is an implicit GEMM exactly like the ConvLSTM cells (convlstm_tma.cu): M = the 128 pixels of a 16 x 8 tile,
// K = 32 channels, and a 3x3 tap is only a different descriptor start address into the staged [18][10] halo tile:
//   t0[p][o] = b[o] + sum_tap sum_c c[p + tap][c] * w[o][c][tap]            (9 taps x 2 K-steps x 2 MMAs per fill, N = 16)
// B rows 0-3 hold fp16 hi(w), rows 8-11 lo(w); A_hi x B and A_lo x B accumulate hi*hi + lo*hi in columns 0-3 and
// hi*lo + lo*lo in columns 8-11 (fp32 in TMEM), and the epilogue adds the two blocks.  Operands: un-swizzled K-major
// fp16 planes [8-channel chunk][row][16 B] (tc_common.cuh).  (The first version computed z[p][tap][o] for all 180 halo
// pixels and summed the shifted taps through shared memory: 430 of its 2 700 L1 wavefronts per fill.)
//
// Work unit = one "fill" = (tile of 16 x 8 pixels + halo, source view, plane); fills run in (tile, view, plane) order so
// a tile's source lines stay in L1 across the planes of a batch.  The CTA (one per SM, persistent) is a ring of
// independent WORKER GROUPS of four warps plus one MMA-issuing warp.  Group g owns operand stage g, recipe slot g and TMEM
// accumulator g and walks fills g, g+G, g+2G, ...:
//   1. one thread per staged pixel evaluates the homography and writes a ready-to-use gather recipe (shared memory);
//   2. 8 lanes per pixel gather branch-free (out-of-image taps read line 0 with weight 0), (warp-ref)^2, fp16 split,
//      stage -> arrive on full[g];                      [the MMA warp issues the 36 MMAs of that fill, commit -> done[g]]
//   3. wait done[g]: TMEM -> t0 (+ bias), GroupNorm-1 sums.
// The groups are out of phase, so one group's load / MMA latency is covered by the others' arithmetic, and no role
// waits on another group's progress (the first version had a single epilogue group that serialised the pipeline).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/aarmvs.h"

namespace {

#ifndef STEM_NBLK
#define STEM_NBLK 1             // 8-column MMA blocks per tile: the tile is 16 x (8 NBLK) pixels
#endif
#ifndef STEM_GROUPS
#define STEM_GROUPS (STEM_NBLK == 1 ? 5 : 3)
#endif
#ifndef STEM_NB
#define STEM_NB 2
#endif
#ifndef STEM_VREUSE
#define STEM_VREUSE 1           // gather threads walk columns of six staged pixels and keep the source lines shared by vertically
                                // adjacent pixels in registers: 14 instead of 24 line reads per column.  (Measured slower before the
                                // recipes moved off the critical path, 0.911 vs 0.873 ms/plane; with them 0.881 vs 0.932.)
#endif
constexpr int NBLK = STEM_NBLK;
// NBLK = 2 (16 x 16 tiles: 324 staged pixels for 256 outputs = 27 % halo instead of 41 %) was measured SLOWER at cfg2
// (0.91 vs 0.86 ms/plane): only three 256-thread groups fit in shared memory and each waits 4 800 cycles per fill for
// its 72 MMAs with too few other groups to cover it.
constexpr int TR = 16, TCOL = 8 * NBLK, HR = TR + 2, HC = TCOL + 2, NPX = HR * HC;
#ifndef STEM_AROWS
#define STEM_AROWS (STEM_NBLK == 1 ? 186 : 330)
#endif
constexpr int AROWS = STEM_AROWS;          // operand rows per K chunk: >= NPX; = 2 mod 8 puts the four chunks a half-warp's
                                           // 8-byte stores touch (2 pixels x 8 lanes) in four disjoint groups of 8 banks
static_assert(AROWS >= NPX && AROWS % 8 == 2, "a chunk holds the halo tile; conflict-free chunk stride");
constexpr int G = STEM_GROUPS;             // worker groups = operand stages = TMEM accumulators
constexpr int GROUP_THREADS = 128 * NBLK;      // one thread per output pixel = per accumulator row
constexpr int N_THREADS = 32 + G * GROUP_THREADS;
constexpr int KC8 = 4, NB = 16;                                          // K chunks; accumulator columns (0-3 and 8-11 used)
constexpr int A_PLANE = KC8 * AROWS * 16, A_STAGE = 2 * A_PLANE;
constexpr int B_BYTES = 9 * KC8 * NB * 16;                               // [tap][chunk][16 rows][16 B] = 9216 bytes
constexpr int TAP_BYTES = NPX * 32;                                      // one TapRec (32 B) per staged pixel
#ifndef STEM_PREFETCH
#define STEM_PREFETCH 1         // recipe threads prefetch the source lines of their pixel (west lines only; 2: all four)
#endif
#ifndef STEM_AHEAD
#define STEM_AHEAD 1            // recipes of a group's NEXT fill are evaluated while the MMAs of the current one run (two recipe slots)
#endif
constexpr int TAP_SLOTS = STEM_AHEAD ? 2 : 1;
constexpr int SMEM_FIXED = B_BYTES + G * (A_STAGE + TAP_SLOTS * TAP_BYTES) + 1024;    // + G * 4 warps * planes * S * 8 bytes of GroupNorm partial sums
#ifndef STEM_STAT_SHARED
#define STEM_STAT_SHARED 1      // 1: one GroupNorm partial-sum slot per worker GROUP (shared-memory float atomics of its four warps)
#endif                          //    instead of one per warp: 3/4 less shared memory for the sums (what a sixth group needs)
constexpr int STAT_ROWS = STEM_STAT_SHARED ? STEM_GROUPS : STEM_GROUPS * 4 * STEM_NBLK;
#ifndef STEM_ACC_STRIDE
#define STEM_ACC_STRIDE 32
#endif
constexpr int ACC = STEM_ACC_STRIDE;       // TMEM columns between the groups' accumulators
constexpr int NACC = G * NBLK;             // accumulators: one per group and 8-column block
constexpr int TMEM_COLS = NACC * ACC <= 32 ? 32 : NACC * ACC <= 64 ? 64 : NACC * ACC <= 128 ? 128 : NACC * ACC <= 256 ? 256 : 512;
static_assert(NACC * ACC <= 512 && ACC >= NB, "accumulators exceed the TMEM allocation");

struct StemArgs {
    const float* ref;
    const float* src;
    size_t view_stride;
    int H, W, S, K;             // K = planes per launch
    const float* cams;
    const __half* weight;       // [9 taps][4 chunks][16 rows: 0-3 hi(w[o]), 8-11 lo(w[o])][8] fp16
    float bias[4];
    const float* depth_values;
    int d_host;
    const int* d_dev;
    int align_corners;
    float* t0;
    double* stats;
    float* costv;               // optional [K][S][P][32] (warp-ref)^2: written for the tile's interior pixels (mode 1) so that pass D
    int costv_mode;             //   need not gather again
#ifdef AARMVS_DEBUG
    unsigned long long* prof;   // debug: clock64 stamps of group 0 of CTA 0, [64 fills][8]
    int dbg;                    // debug ablation flags (aarmvs_debug_flags): 2 skip recipes, 4 skip MMA
#endif
};

// Debug hooks (timing experiments of scripts/stem_ablation.py) exist only in -DAARMVS_DEBUG builds: the product library has
// neither the process-global switches nor their branches in the hot loops.
#ifdef AARMVS_DEBUG
#define STEM_DBG(a, flag) ((a).dbg & (flag))
#define STEM_PROF(a) ((a).prof)
#else
#define STEM_DBG(a, flag) (0)
#define STEM_PROF(a) ((unsigned long long*)nullptr)
#endif

// Gather recipe of one staged pixel (32 bytes; the L1 data pipe bounds this kernel, so the record is kept to two loads):
// the four bilinear weights and the BYTE offset of the north-west line (common.cuh: make_recipe -- fixed +128 / +row steps,
// weight 0 for a corner outside the image), and the byte offset of the reference pixel with the "this fill owns the pixel's
// costv line" flag in bit 0 (-1: the pixel lies outside the image, its cost is the conv's zero padding).
struct __align__(16) TapRec {
    float w[4];
    unsigned base;
    int ref;
    unsigned pad0, pad1;
};

__device__ __forceinline__ float4 ldg4b(const unsigned char* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void sts64(uint32_t addr, uint2 v) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

// KEEP_CV: also store (warp - ref)^2 of the tile's interior pixels to a.costv (the training backward's per-plane recompute and
// Engine(keep_costv=True)); a template parameter so that the default kernel carries none of its predicated-off instructions.
template <bool KEEP_CV>
__global__ void __launch_bounds__(N_THREADS, 1) cost_stem_tc_kernel(const __grid_constant__ StemArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;
    unsigned char* sA = smem + B_BYTES;
    unsigned char* sT = sA + G * A_STAGE;
    const int SK = a.S * a.K;
    float2* sstat = reinterpret_cast<float2*>(sT + G * TAP_SLOTS * TAP_BYTES);      // [worker warp][(plane, view)] GroupNorm partial sums: no atomics
    __shared__ uint64_t full_a[G], done_a[G];
    __shared__ uint32_t tmem_base_s;
    __shared__ float scam[AARMVS_MAX_SRC * 12];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* depths = a.depth_values + plane_index(a.d_host, a.d_dev);

    for (int i = tid; i < B_BYTES / 16; i += N_THREADS)
        reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(a.weight) + i);
    for (int i = tid; i < a.S * 12; i += N_THREADS) scam[i] = a.cams[i];
    for (int i = tid; i < STAT_ROWS * SK; i += N_THREADS) sstat[i] = make_float2(0.f, 0.f);
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
    const int total = my_tiles * SK;

    if (warp == 0) {
        // =========================================== MMA issuer ===========================================
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::idesc_f16_f32(128, NB);
        const uint32_t a_base = tc::smem_u32(sA), b_base = tc::smem_u32(sB);
        constexpr uint32_t A_HI = tc::desc_hi(HC * 16), B_HI = tc::desc_hi(8 * 16);    // 8-row groups: one tile row / 8 B rows apart
        const uint32_t b_lo = tc::desc_lo(b_base, NB * 16);
        for (int it = 0; it < total; ++it) {
            const int g = it % G;
            tc::mbar_wait(&full_a[g], (it / G) & 1);
            tc::tc_fence_after();
            const uint32_t a_lo_hi = tc::desc_lo(a_base + g * A_STAGE, AROWS * 16);
            const uint32_t a_lo_lo = tc::desc_lo(a_base + g * A_STAGE + A_PLANE, AROWS * 16);
            if (leader) {
                if (!STEM_DBG(a, 4))
#pragma unroll
                for (int mb = 0; mb < NBLK; ++mb) {          // 8-column block of the tile: 8 staged pixels further
                    const uint32_t dst = tmem_base + (g * NBLK + mb) * ACC;
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint32_t ak = (uint32_t)((2 * ks) * AROWS + (tap / 3) * HC + (tap % 3) + mb * 8);      // 16-byte units
                            const uint32_t bk = (uint32_t)((tap * KC8 + 2 * ks) * NB);
                            tc::mma_f16_parts(dst, a_lo_hi + ak, A_HI, b_lo + bk, B_HI, idesc, (tap | ks) != 0);
                            tc::mma_f16_parts(dst, a_lo_lo + ak, A_HI, b_lo + bk, B_HI, idesc, 1);
                        }
                    }
                }
                tc::mma_commit(&done_a[g]);        // operand stage free AND accumulator complete
            }
            __syncwarp();
        }
    } else {
        // =========================================== worker groups ===========================================
        const int g = (warp - 1) / (4 * NBLK), lt = tid - 32 - g * GROUP_THREADS;
        const int mb = ((warp - 1) >> 2) % NBLK;                      // 8-column block whose accumulator this warp reads
        const int quad = warp & 3;                                   // TMEM lane quadrant this warp may read
        unsigned char* stage = sA + g * A_STAGE;
        TapRec* const taps_g = reinterpret_cast<TapRec*>(sT + g * TAP_SLOTS * TAP_BYTES);
        const uint32_t bar_id = 1 + g;
        constexpr int ITEMS = (NPX * 8 + GROUP_THREADS - 1) / GROUP_THREADS;     // 12 (the last one partial)
        constexpr int NB = STEM_NB;
        // a thread always handles the same 16-byte slice (4 channels) of its pixels: item j is pixel px0 + (GROUP_THREADS / 8) j
        const int c4i = lt & 7, px0 = lt >> 3;
        const uint32_t st_thread = tc::smem_u32(stage) + (uint32_t)((c4i >> 1) * AROWS * 16 + (c4i & 1) * 8);
        constexpr bool keep_cv = KEEP_CV;
        const unsigned row_bytes = (unsigned)a.W * 128u;
        // fill it = g, g + G, ... decoded incrementally as (tile slot tl, view v, plane k): it = (tl * S + v) * K + k
        int tl = g / SK, v = (g % SK) / a.K, k = g % a.K, n = 0;
        // gather recipes of fill (tile slot tl_, view v_, plane k_): one thread per staged pixel
        auto write_recipes = [&](int ty_, int tx_, int v_, int k_, TapRec* dst) {
            const float depth_ = depths[k_];
            for (int px = lt; px < NPX; px += GROUP_THREADS) {
                const int r = px / HC, cc = px - r * HC;
                const int y = ty_ * TR - 1 + r, x = tx_ * TCOL - 1 + cc;
                TapRec rec;
                rec.w[0] = rec.w[1] = rec.w[2] = rec.w[3] = 0.f;
                rec.base = 0u; rec.ref = -1; rec.pad0 = rec.pad1 = 0u;
                if (y >= 0 && y < a.H && x >= 0 && x < a.W && !STEM_DBG(a, 2)) {
                    float4 wts;
                    make_recipe(scam + v_ * 12, x, y, depth_, a.W, a.H, a.align_corners != 0, wts, rec.base);
                    rec.w[0] = wts.x; rec.w[1] = wts.y; rec.w[2] = wts.z; rec.w[3] = wts.w;
#if STEM_PREFETCH
                    {   // the gather of this fill starts thousands of cycles from now (after the current fill's MMAs): pull its
                        // source lines towards L1 meanwhile (ncu: 40 % of the warp stalls are the gather's first use of a load)
                        const unsigned char* sb = reinterpret_cast<const unsigned char*>(a.src + (size_t)v_ * a.view_stride);
                        const unsigned o0 = rec.base, ox = 128u, od = (unsigned)a.W * 128u;
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(sb + o0));
                        if (od) asm volatile("prefetch.global.L1 [%0];" ::"l"(sb + (o0 + od)));
                        if (ox && (cc == HC - 1 || STEM_PREFETCH > 1)) {      // the east lines are the next pixel's west lines, except in the last column
                            asm volatile("prefetch.global.L1 [%0];" ::"l"(sb + (o0 + ox)));
                            if (od) asm volatile("prefetch.global.L1 [%0];" ::"l"(sb + (o0 + ox + od)));
                        }
                    }
#endif
                    rec.ref = (y * a.W + x) * 128 + ((r >= 1 && r <= TR && cc >= 1 && cc <= TCOL) ? 1 : 0);
                }
                dst[px] = rec;
            }
        };
        // tile coordinates of the current fill: the division is redone only when a group moves on to another tile
        int ty = (int)(blockIdx.x + tl * gridDim.x) / tiles_x, tx = (int)(blockIdx.x + tl * gridDim.x) - ty * tiles_x;
        if (STEM_AHEAD && g < total) write_recipes(ty, tx, v, k, taps_g);
        for (int it = g; it < total; it += G, ++n) {
            const bool pf = STEM_PROF(a) && blockIdx.x == 0 && g == 0 && lt == 0 && n < 64;
            unsigned long long* pp = STEM_PROF(a) + n * 8;
            if (pf) pp[0] = clock64();
            // ---- 1. gather recipes
            TapRec* const taps = taps_g + (STEM_AHEAD ? (n & 1) * NPX : 0);
            if (!STEM_AHEAD) write_recipes(ty, tx, v, k, taps);
            // (the stage and the recipes of the previous fill are free: every thread of the group arrived on full[g] after its
            //  last read, and done[g] was observed after the MMAs' last read)
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(GROUP_THREADS) : "memory");
            if (pf) pp[1] = clock64();
            // ---- 2. gather, (warp - ref)^2, split, stage
            const float* sv = a.src + (size_t)v * a.view_stride;
            float* costv_slot = KEEP_CV ? a.costv + ((size_t)(k * a.S + v) * P) * 32 : nullptr;
            {
                const unsigned char* svb = reinterpret_cast<const unsigned char*>(sv) + c4i * 16;
                const unsigned char* refb = reinterpret_cast<const unsigned char*>(a.ref) + c4i * 16;
                unsigned char* cvb = reinterpret_cast<unsigned char*>(costv_slot) + c4i * 16;
                // one staged pixel: blend in ATen's order (see gather_taps): fma(se, w, fma(sw, w, fma(ne, w, nw * w))), two channels
                // per instruction; (warp - ref)^2; costv line; fp16 split; stage
                auto finish = [&](const float4& ta, const float4& tb, const float4& tcn, const float4& td, const float4& rv,
                                  const float4& wq, int rk, int pxr, const bool in_tile) {
                    const uint64_t w0 = pk2(wq.x, wq.x), w1 = pk2(wq.y, wq.y), w2 = pk2(wq.z, wq.z), w3 = pk2(wq.w, wq.w);
                    const uint64_t g01 = fma2(pk2(td.x, td.y), w3, fma2(pk2(tcn.x, tcn.y), w2, fma2(pk2(tb.x, tb.y), w1, mul2(pk2(ta.x, ta.y), w0))));
                    const uint64_t g23 = fma2(pk2(td.z, td.w), w3, fma2(pk2(tcn.z, tcn.w), w2, fma2(pk2(tb.z, tb.w), w1, mul2(pk2(ta.z, ta.w), w0))));
                    // outside the image the cost is the conv's zero padding: d * 0 = +-0 exactly (finite features), d * 1 = d
                    const float inf_ = rk >= 0 ? 1.0f : 0.0f;
                    const uint64_t in2 = pk2(inf_, inf_);
                    const uint64_t d01 = mul2(sub2(g01, pk2(rv.x, rv.y)), in2), d23 = mul2(sub2(g23, pk2(rv.z, rv.w)), in2);
                    const uint64_t c01 = mul2(d01, d01), c23 = mul2(d23, d23);
                    float4 cv;
                    upk2(c01, cv.x, cv.y);
                    upk2(c23, cv.z, cv.w);
                    if (keep_cv && rk >= 0 && (rk & 1) && in_tile)          // this fill owns the pixel's costv line
                        *reinterpret_cast<float4*>(cvb + ((unsigned)rk & ~127u)) = cv;
                    // fp16 split as tc::split4_f16: hi = top 11 mantissa bits, lo = fp16(x - hi)
                    const float h0 = __uint_as_float(__float_as_uint(cv.x) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(cv.y) & 0xFFFFE000u);
                    const float h2 = __uint_as_float(__float_as_uint(cv.z) & 0xFFFFE000u), h3 = __uint_as_float(__float_as_uint(cv.w) & 0xFFFFE000u);
                    float l0, l1, l2, l3;
                    upk2(sub2(c01, pk2(h0, h1)), l0, l1);
                    upk2(sub2(c23, pk2(h2, h3)), l2, l3);
                    const __half2 ha = __floats2half2_rn(h0, h1), hb = __floats2half2_rn(h2, h3);
                    const __half2 la = __floats2half2_rn(l0, l1), lb = __floats2half2_rn(l2, l3);
                    const uint2 hi = make_uint2(*reinterpret_cast<const uint32_t*>(&ha), *reinterpret_cast<const uint32_t*>(&hb));
                    const uint2 lo = make_uint2(*reinterpret_cast<const uint32_t*>(&la), *reinterpret_cast<const uint32_t*>(&lb));
                    if (in_tile) {
                        sts64(st_thread + pxr * 16, hi);
                        sts64(st_thread + A_PLANE + pxr * 16, lo);
                    }
                };
#if STEM_VREUSE
                // A thread walks UROWS vertically adjacent staged pixels: where the warp is locally a translation, the south
                // lines of one pixel ARE the north lines of the pixel below (same addresses), so they are taken from registers
                // instead of being gathered again -- 14 instead of 24 line reads per unit.  The L1 data pipe (one wavefront per
                // 128-byte line) is what bounds this kernel.
                constexpr int UROWS = 6, UNITS = (HR / UROWS) * HC;
                static_assert(HR % UROWS == 0 && UROWS % 2 == 0, "units tile the staged rows");
#pragma unroll 1
                for (int u = px0; u < UNITS; u += GROUP_THREADS / 8) {
                    const int ut = u / HC;
                    int pxr = ut * (UROWS - 1) * HC + u;                  // (row ut * UROWS, column u - ut * HC)
                    float4 pc = make_float4(0.f, 0.f, 0.f, 0.f), pd = pc;
                    unsigned want = 1u;                                   // north-west line offset the registers hold (odd = none: offsets are multiples of 128)
#pragma unroll 1
                    for (int jb = 0; jb < UROWS; jb += 2, pxr += 2 * HC) {
                        // (written out for two rows: all loads of both rows are issued before the first use; a generic
                        //  array-and-loop form of this body was scheduled with row 0's arithmetic ahead of row 1's loads and
                        //  measured 0.957 instead of 0.881 ms/plane)
                        const TapRec* rec0 = taps + pxr;
                        const TapRec* rec1 = rec0 + HC;
                        const float4 wq0 = *reinterpret_cast<const float4*>(rec0->w), wq1 = *reinterpret_cast<const float4*>(rec1->w);
                        const uint2 br0 = *reinterpret_cast<const uint2*>(&rec0->base), br1 = *reinterpret_cast<const uint2*>(&rec1->base);
                        // the four lines of a pixel: north-west at svb + base, the others at fixed steps (+128, +row, +row + 128):
                        // one 64-bit add per row, the column step is an immediate of the load
                        const unsigned char* n0 = svb + br0.x;
                        const unsigned char* s0 = n0 + row_bytes;
                        const unsigned char* n1 = svb + br1.x;
                        const unsigned char* s1 = n1 + row_bytes;
                        // a pixel's north lines = the south lines of the pixel above: taken from registers
                        const bool ra = br0.x == want;
                        const bool rb = br1.x == br0.x + row_bytes;
                        float4 ta0 = pc, tb0 = pd, ta1, tb1;
                        if (!ra) { ta0 = ldg4b(n0); tb0 = ldg4b(n0 + 128); }
                        const float4 tc0 = ldg4b(s0), td0 = ldg4b(s0 + 128);
                        if (!rb) { ta1 = ldg4b(n1); tb1 = ldg4b(n1 + 128); }
                        const float4 tc1 = ldg4b(s1), td1 = ldg4b(s1 + 128);
                        const float4 rv0 = ldg4b(refb + ((unsigned)max((int)br0.y, 0) & ~127u));
                        const float4 rv1 = ldg4b(refb + ((unsigned)max((int)br1.y, 0) & ~127u));
                        if (rb) { ta1 = tc0; tb1 = td0; }
                        finish(ta0, tb0, tc0, td0, rv0, wq0, (int)br0.y, pxr, true);          // units never leave the staged tile
                        finish(ta1, tb1, tc1, td1, rv1, wq1, (int)br1.y, pxr + HC, true);
                        pc = tc1; pd = td1;
                        want = br1.x + row_bytes;
                    }
                }
#else
#pragma unroll 1
                for (int jb = 0; jb < ITEMS; jb += NB) {
                    float4 ta[NB], tb[NB], tcn[NB], td[NB], rv[NB], wq[NB];
                    int rk[NB];
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        const TapRec* rec = taps + min(px0 + (jb + j) * (GROUP_THREADS / 8), NPX - 1);
                        wq[j] = *reinterpret_cast<const float4*>(rec->w);
                        const uint2 br = *reinterpret_cast<const uint2*>(&rec->base);
                        rk[j] = (int)br.y;
                        const unsigned o0 = br.x, o1 = o0 + 128u, od = row_bytes;
                        ta[j] = ldg4b(svb + o0);
                        tb[j] = ldg4b(svb + o1);
                        tcn[j] = ldg4b(svb + (o0 + od));
                        td[j] = ldg4b(svb + (o1 + od));
                        rv[j] = ldg4b(refb + ((unsigned)max(rk[j], 0) & ~127u));
                    }
#pragma unroll
                    for (int j = 0; j < NB; ++j)
                        finish(ta[j], tb[j], tcn[j], td[j], rv[j], wq[j], rk[j], px0 + (jb + j) * (GROUP_THREADS / 8),
                               px0 + (jb + j) * (GROUP_THREADS / 8) < NPX);
                }
#endif
            }
            if (pf) pp[2] = clock64();
            tc::fence_async_smem();
            tc::mbar_arrive(&full_a[g]);
            // next fill of this group: (tl2, v2, k2).  Its recipes go to the other slot while this fill's MMAs run: every thread of
            // the group has passed this fill's bar.sync, i.e. finished reading that slot in the previous fill.
            int tl2 = tl, v2 = v, k2 = k + G;
            while (k2 >= a.K) { k2 -= a.K; if (++v2 == a.S) { v2 = 0; ++tl2; } }
            int ty2 = ty, tx2 = tx;
            if (tl2 != tl) {
                const int tile2 = blockIdx.x + tl2 * gridDim.x;
                ty2 = tile2 / tiles_x;
                tx2 = tile2 - ty2 * tiles_x;
            }
            if (STEM_AHEAD && it + G < total) write_recipes(ty2, tx2, v2, k2, taps_g + ((n + 1) & 1) * NPX);
            if (pf) pp[3] = clock64();
            // ---- 3. accumulator row of this thread's pixel -> t0
            tc::mbar_wait_relaxed(&done_a[g], n & 1);
            tc::tc_fence_after();
            if (pf) pp[4] = clock64();
            float4 o;
            {
                uint32_t r[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (g * NBLK + mb) * ACC;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                o = make_float4(a.bias[0] + (__uint_as_float(r[0]) + __uint_as_float(r[8])), a.bias[1] + (__uint_as_float(r[1]) + __uint_as_float(r[9])),
                                a.bias[2] + (__uint_as_float(r[2]) + __uint_as_float(r[10])), a.bias[3] + (__uint_as_float(r[3]) + __uint_as_float(r[11])));
            }
            tc::tc_fence_before();
            if (pf) pp[5] = clock64();
            const int m = quad * 32 + lane;                             // accumulator row = pixel (m >> 3, m & 7) of the tile
            const int y = ty * TR + (m >> 3), x = tx * TCOL + mb * 8 + (m & 7);
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
#if STEM_STAT_SHARED
                atomicAdd(&sstat[g * SK + slot].x, s);
                atomicAdd(&sstat[g * SK + slot].y, ss);
#else
                float2 acc = sstat[(warp - 1) * SK + slot];
                acc.x += s; acc.y += ss;
                sstat[(warp - 1) * SK + slot] = acc;
#endif
            }
            if (pf) pp[6] = clock64();
            tl = tl2; v = v2; k = k2; ty = ty2; tx = tx2;
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<TMEM_COLS>(tmem_base);
    if (tid < a.K * a.S) {                                                   // stats[k][v][GN1][sum, sumsq]
        double s = 0.0, ss = 0.0;
        for (int w = 0; w < STAT_ROWS; ++w) { s += (double)sstat[w * SK + tid].x; ss += (double)sstat[w * SK + tid].y; }
        atomicAdd(a.stats + tid * 6 + 0, s);
        atomicAdd(a.stats + tid * 6 + 1, ss);
    }
}

#ifdef AARMVS_DEBUG
int g_dbg = 0;
unsigned long long* g_stem_prof = nullptr;
#endif

}  // namespace

#ifdef AARMVS_DEBUG
extern "C" void aarmvs_debug_flags(int flags) { g_dbg = flags; }
extern "C" void aarmvs_debug_stem_profile_buffer(void* buf) { g_stem_prof = (unsigned long long*)buf; }
#endif

extern "C" int aarmvs_cost_stem_tc_max_slots(void) { return (227 * 1024 - 2048 - SMEM_FIXED) / (STAT_ROWS * 8); }

namespace {
int launch_cost_stem_tc(const float* ref, const float* src, size_t view_stride, int H, int W, int S, const float* cams,
                        const void* stem_w_f16, const float* stem_b_host, const float* depth_values, int d_host, const int* d_dev,
                        int planes, int align_corners, float* t0, double* stats, float* costv, int costv_mode, void* stream) {
    AARMVS_CHECK_ARG(t0 && stats && stem_w_f16 && stem_b_host && depth_values && H > 0 && W > 0);
    AARMVS_CHECK_ARG(costv_mode >= 0 && costv_mode <= 1 && (costv_mode == 0 || costv));
    AARMVS_CHECK_ARG((size_t)H * W * 128 < 0x7FFFFFFFull);          // gather recipes hold 32-bit byte offsets (TapRec.ref is signed)
    AARMVS_CHECK_ARG(ref && src && cams);
    AARMVS_CHECK_ARG(S >= 1 && S <= AARMVS_MAX_SRC && planes >= 1 && planes <= AARMVS_MAX_PLANES);
    StemArgs a = {};
    a.K = planes;
    a.ref = ref; a.src = src; a.view_stride = view_stride; a.H = H; a.W = W; a.S = S; a.cams = cams;
    a.weight = (const __half*)stem_w_f16;
    for (int i = 0; i < 4; ++i) a.bias[i] = stem_b_host[i];
    a.depth_values = depth_values; a.d_host = d_host; a.d_dev = d_dev; a.align_corners = align_corners; a.t0 = t0;
    a.stats = stats;
    a.costv = costv; a.costv_mode = costv_mode;
#ifdef AARMVS_DEBUG
    a.dbg = g_dbg;
    a.prof = g_stem_prof;
#endif
    static unsigned long long smem_ok0 = 0, smem_ok1 = 0;
    constexpr int SMEM_CAP = 227 * 1024 - 2048;      // the kernel also has ~1 KB of static shared memory
    static_assert(SMEM_FIXED + STAT_ROWS * 8 * 8 <= SMEM_CAP, "shared memory budget");
    if (cudaError_t e = costv_mode == 1 ? ensure_dyn_smem(cost_stem_tc_kernel<true>, SMEM_CAP, smem_ok1)
                                        : ensure_dyn_smem(cost_stem_tc_kernel<false>, SMEM_CAP, smem_ok0)) return (int)e;
    const int SMEM = SMEM_FIXED + STAT_ROWS * planes * S * 8;
    AARMVS_CHECK_ARG(SMEM <= SMEM_CAP);          // planes * S <= aarmvs_cost_stem_tc_max_slots()
    const int ntiles = ((W + TCOL - 1) / TCOL) * ((H + TR - 1) / TR);
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    if (costv_mode == 1) cost_stem_tc_kernel<true><<<ntiles < sms ? ntiles : sms, N_THREADS, SMEM, (cudaStream_t)stream>>>(a);
    else cost_stem_tc_kernel<false><<<ntiles < sms ? ntiles : sms, N_THREADS, SMEM, (cudaStream_t)stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}
}  // namespace

extern "C" int aarmvs_cost_stem_tc(const float* ref, const float* src, size_t view_stride, int H, int W, int S,
                                   const float* cams, const void* stem_w_f16, const float* stem_b_host,
                                   const float* depth_values, int d_host, const int* d_dev, int planes, int align_corners,
                                   float* t0, double* stats, float* costv, int costv_mode, void* stream) {
    return launch_cost_stem_tc(ref, src, view_stride, H, W, S, cams, stem_w_f16, stem_b_host, depth_values, d_host, d_dev, planes,
                               align_corners, t0, stats, costv, costv_mode, stream);
}
