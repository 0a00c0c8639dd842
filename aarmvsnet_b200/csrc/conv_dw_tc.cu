// Weight gradient of the ConvLSTM cells on the tensor cores (BASELINE config 5; the backward of models/module.py:80-82):
//
//   dW[tap][ci][co] = sum over pixels p:  X[p + tap][ci] * dG[p][co]          a GEMM whose K dimension is the PIXELS
//
// Operand planes [2 (hi, lo)][C/8][H][W][8 x fp16] (include/aarmvs.h) are K-major operands when K = channels (the forward
// cell, conv_dx_tc.cu) and, read the other way, exactly the canonical un-swizzled MN-MAJOR operand when K = pixels: a 16-byte
// element = 8 consecutive M (or N) indices of one K index, the next K index 16 bytes further, the next group of eight K
// indices (the next row of the staged 18 x 10 halo tile) 160 bytes further (LBO), the next eight channels one chunk further
// (SBO = 2 880 bytes).  So the SAME TMA boxes feed this GEMM with both descriptors flagged MN-major, and a 3x3 tap is again
// only a different start address of the X operand.
//
// One MMA (M = 128, N = 2 * 4*hidden, K = 16 pixels = two tile rows) per tap and K step covers all split-fp16 products at once:
// the 128 M rows run over the chunks of X_hi, then X_lo (the lo plane follows the hi plane in the stage; the remaining rows read
// whatever follows and are ignored -- rows are independent), the N columns over the chunks of dG_hi, then dG_lo:
//   rows [0, cx) x cols [0, cy): hi*hi    rows [lo, lo + cx) x cols [0, cy): lo*hi    rows [0, cx) x cols [cy, 2cy): hi*lo
// A CTA owns one kernel row ky: three taps = three accumulators of N columns in TMEM, which stay there over ALL its tiles
// (accumulate flag), so there is one epilogue per CTA: TMEM -> (lo*hi rows added to the hi rows) -> fp32 atomics into dW, times 2^-k of the gate gradients' scale
// (conv_dx_tc.cu).  X is the channel concat [x | h_{d-1}] of the cell's inputs, split into planes by aarmvs_split_planes_cat.
#include "common.cuh"
#include "tc_common.cuh"
#include "plane_map.cuh"
#include "../../include/aarmvs.h"

namespace {

constexpr int TR = 16, TCOL = 8, HR = TR + 2, HC = TCOL + 2, NPX = HR * HC;   // 180 staged pixels per tile
constexpr int CHUNK_BYTES = NPX * 16;
static_assert(HR == planemap::BOX_ROWS && HC == planemap::BOX_COLS, "the tensor-map box is the staged halo tile");
constexpr int N_EPI_WARPS = 16, MMA_WARP = N_EPI_WARPS, TMA_WARP = MMA_WARP + 1;
constexpr int N_THREADS = (N_EPI_WARPS + 2) * 32;                             // 576
constexpr int DW_GROUPS = 8;          // measured at cfg5: 4..8 groups 213 ms per step, 16: 217, 32: 228

struct alignas(64) DwArgs {
    CUtensorMap xmap, gmap;     // X and gate-gradient operand planes, box (80, 18, chunks, 1, 1)
    int H, W, groups;           // groups: tile groups (gridDim.x = 3 * groups)
    const float* inv_scale;     // device: 2^-k of the gate-gradient planes' scale
    float* dw;                  // [9][cx][cy], +=
};

// XC8 / GC8: 8-channel chunks of X / dG; XS8: chunks one X plane occupies in the stage (cell 4's five chunks are padded to six:
// a chunk is 2 880 bytes and TMA destinations must be 128-byte aligned; the pad rows are ignored like the trailing ones)
template <int XC8, int XS8, int GC8, int NST>
struct Cfg {
    static constexpr int CX = XC8 * 8, LO_ROW = XS8 * 8, CY = GC8 * 8, N = 2 * CY;
    static constexpr int X_BYTES = 2 * XS8 * CHUNK_BYTES, G_BYTES = 2 * GC8 * CHUNK_BYTES;
    static constexpr int STAGE = X_BYTES + G_BYTES;
    static constexpr int TX_BYTES = (2 * XC8 + 2 * GC8) * CHUNK_BYTES;        // what the four tensor copies of a tile deliver
    static constexpr int SMEM = NST * STAGE + 1024;
    static constexpr int TMEM_COLS = 3 * N <= 256 ? 256 : 512;
    static_assert(XS8 >= XC8 && 2 * XS8 + 2 * GC8 >= 16, "the 128 M rows (16 chunks from the start of X_hi) must stay inside the stage");
    static_assert(2 * XS8 <= 16 && N <= 256 && N % 16 == 0, "operand shape");
    static_assert(STAGE % 128 == 0 && (X_BYTES / 2) % 128 == 0 && (G_BYTES / 2) % 128 == 0, "TMA destinations must be 128-byte aligned");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}

template <int XC8, int XS8, int GC8, int NST>
__global__ void __launch_bounds__(N_THREADS, 1) conv_dw_tc_kernel(const __grid_constant__ DwArgs a) {
    using C = Cfg<XC8, XS8, GC8, NST>;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full_a[NST], empty_a[NST], done_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ky = blockIdx.x % 3, group = blockIdx.x / 3;
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { tc::mbar_init(&full_a[s], 1); tc::mbar_init(&empty_a[s], 1); }
        tc::mbar_init(&done_bar, 1);
        tc::mbar_fence_init();
    }
    if (warp == MMA_WARP) tc::tmem_alloc<C::TMEM_COLS>(&tmem_base_s);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const int tiles_x = (a.W + TCOL - 1) / TCOL, tiles_y = (a.H + TR - 1) / TR;
    const int ntiles = tiles_x * tiles_y;

    if (warp == TMA_WARP) {
        // =========================================== TMA producer ===========================================
        if (lane == 0) { tc::tma_prefetch_desc(&a.xmap); tc::tma_prefetch_desc(&a.gmap); }
        int it = 0;
        for (int tile = group; tile < ntiles; tile += a.groups, ++it) {
            const int st = it % NST;
            if (lane == 0) {
                tc::mbar_wait(&empty_a[st], ((it / NST) & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&full_a[st], C::TX_BYTES);
            }
            __syncwarp();
            if (lane < 4) {                              // X hi, X lo, dG hi, dG lo: one warp instruction
                const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
                const int hl = lane & 1;
                unsigned char* dst = smem + st * C::STAGE + (lane < 2 ? hl * (C::X_BYTES / 2) : C::X_BYTES + hl * (C::G_BYTES / 2));
                tc::tma_load_5d(dst, lane < 2 ? &a.xmap : &a.gmap, &full_a[st], (tx * TCOL - 1) * 8, ty * TR - 1, 0, hl, 0);
            }
            __syncwarp();
        }
    } else if (warp == MMA_WARP) {
        // =========================================== MMA issuer ===========================================
        const bool leader = tc::elect_one();
        constexpr uint32_t MN_MAJOR = (1u << 15) | (1u << 16);                 // both operands are read "transposed"
        const uint32_t idesc = tc::idesc_f16_f32(128, C::N) | MN_MAJOR;
        constexpr uint32_t D_HI = tc::desc_hi(CHUNK_BYTES);                     // SBO: the next eight channels
        const uint32_t s_base = tc::smem_u32(smem);
        int it = 0;
        for (int tile = group; tile < ntiles; tile += a.groups, ++it) {
            const int st = it % NST;
            tc::mbar_wait(&full_a[st], (it / NST) & 1);
            tc::tc_fence_after();
            if (leader) {
                const uint32_t xs = s_base + st * C::STAGE, gs = xs + C::X_BYTES;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
                    for (int k2 = 0; k2 < TR / 2; ++k2) {                       // K = 16 pixels = tile rows 2 k2, 2 k2 + 1
                        const uint32_t ad = tc::desc_lo(xs + ((ky + 2 * k2) * HC + kx) * 16, HC * 16);
                        const uint32_t bd = tc::desc_lo(gs + ((1 + 2 * k2) * HC + 1) * 16, HC * 16);
                        tc::mma_f16_parts(tmem_base + kx * C::N, ad, D_HI, bd, D_HI, idesc, (it | k2) != 0);
                    }
                }
                tc::mma_commit(&empty_a[st]);            // operand stage free once these MMAs have read it
            }
            __syncwarp();
        }
        if (leader) tc::mma_commit(&done_bar);           // all accumulators complete
        __syncwarp();
    } else {
        // =========================================== epilogue (once) ===========================================
        // The lo*hi rows hand their sums to the hi rows through shared memory (the operand stages are free: every MMA is
        // complete), so each dW element leaves the CTA as ONE atomic.
        const int q = warp & 3, part = warp >> 2;        // TMEM lane quadrant, quarter of the output channels
        const int m = q * 32 + lane;                     // M row: [0, CX) = X_hi channel m, [LO_ROW, LO_ROW + CX) = X_lo channel m - LO_ROW
        constexpr int NCO = C::CY / 4;                   // 16 or 8 output channels per thread and tap
        tc::mbar_wait_relaxed(&done_bar, 0);
        tc::tc_fence_after();
        const float inv = *a.inv_scale;
        const bool hi_row = m < C::CX, lo_row = m >= C::LO_ROW && m < C::LO_ROW + C::CX;
        const int ci = hi_row ? m : m - C::LO_ROW;
        float* xch = reinterpret_cast<float*>(smem);     // [3 taps][CY][CX]: conflict-free for a warp of consecutive rows
        float acc[3][NCO];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
#pragma unroll
            for (int j = 0; j < NCO; j += 8) {
                uint32_t g[8], g2[8];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kx * C::N + part * NCO + j;
                tmem_ld8_nowait(taddr, g);
                tmem_ld8_nowait(taddr + C::CY, g2);      // the hi*lo block (only the hi rows use it)
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[kx][j + e] = hi_row ? __uint_as_float(g[e]) + __uint_as_float(g2[e]) : __uint_as_float(g[e]);
            }
        }
        if (lo_row) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int e = 0; e < NCO; ++e) xch[(kx * C::CY + part * NCO + e) * C::CX + ci] = acc[kx][e];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_WARPS * 32) : "memory");      // the epilogue warps only
        if (hi_row) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                float* p = a.dw + ((size_t)(ky * 3 + kx) * C::CX + ci) * C::CY + part * NCO;
#pragma unroll
                for (int e = 0; e < NCO; ++e)
                    atomicAdd(p + e, (acc[kx][e] + xch[(kx * C::CY + part * NCO + e) * C::CX + ci]) * inv);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int XC8, int XS8, int GC8, int NST>
int launch_dw(DwArgs& a, cudaStream_t stream) {
    using C = Cfg<XC8, XS8, GC8, NST>;
    auto kern = conv_dw_tc_kernel<XC8, XS8, GC8, NST>;
    static unsigned long long smem_ok = 0;
    if (cudaError_t e = ensure_dyn_smem(kern, C::SMEM, smem_ok)) return (int)e;
    const int ntiles = ((a.W + TCOL - 1) / TCOL) * ((a.H + TR - 1) / TR);
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    // three CTAs (kernel rows) per tile group.  Few groups: the atomics and the per-CTA set-up shrink with the CTA count, the
    // kernel runs beside the dependent chain of the plane (it needs a third of the SMs), and a CTA's MMAs hide its tile loads
    (void)sms;
    a.groups = ntiles < DW_GROUPS ? ntiles : DW_GROUPS;
    kern<<<3 * a.groups, N_THREADS, C::SMEM, stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}

// Channel concat of up to three fp32 maps [P][ch_s] -> operand planes [2][C/8][P][8] (unscaled: activations are O(1)).
struct CatArgs {
    const float* ptr[3];
    int ch[3], n;
};
__global__ void __launch_bounds__(256) split_planes_cat_kernel(const CatArgs a, size_t P, int C, __half* __restrict__ planes) {
    const int chunks = C / 8;
    const size_t n = P * chunks;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / chunks;
        const int c = (int)(i - p * chunks);
        int base = 0, s = 0;
        while (s < a.n - 1 && c * 8 >= base + a.ch[s]) { base += a.ch[s]; ++s; }
        const float* xp = a.ptr[s] + p * a.ch[s] + (c * 8 - base);
        uint2 h0, l0, h1, l1;
        tc::split4_f16(ld4(xp), h0, l0);
        tc::split4_f16(ld4(xp + 4), h1, l1);
        *reinterpret_cast<uint4*>(planes + ((size_t)c * P + p) * 8) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4*>(planes + ((size_t)(chunks + c) * P + p) * 8) = make_uint4(l0.x, l0.y, l1.x, l1.y);
    }
}

}  // namespace

extern "C" int aarmvs_split_planes_cat(const aarmvs_seg* x, int nx, int H, int W, void* planes, void* stream) {
    AARMVS_CHECK_ARG(x && planes && nx >= 1 && nx <= 3 && H > 0 && W > 0 && ((uintptr_t)planes & 15) == 0);
    CatArgs a = {};
    int C = 0;
    for (int s = 0; s < nx; ++s) {
        AARMVS_CHECK_ARG(x[s].ptr && x[s].channels > 0 && x[s].channels % 8 == 0 && ((uintptr_t)x[s].ptr & 15) == 0);
        a.ptr[s] = x[s].ptr; a.ch[s] = x[s].channels;
        C += x[s].channels;
    }
    a.n = nx;
    const size_t P = (size_t)H * W;
    const size_t blocks = (P * (C / 8) + 255) / 256;
    split_planes_cat_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(a, P, C, (__half*)planes);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_conv_dw_tc(const void* x_planes, const void* dgate_planes, int H, int W, int cx, int cy, const float* inv_scale,
                                 float* dweight, void* stream) {
    AARMVS_CHECK_ARG(x_planes && dgate_planes && inv_scale && dweight && H > 0 && W > 0);
    AARMVS_CHECK_ARG(((uintptr_t)x_planes & 15) == 0 && ((uintptr_t)dgate_planes & 15) == 0);
    DwArgs a = {};
    if (int e = planemap::make_plane_map(&a.xmap, x_planes, H, W, cx / 8, 1)) return e;
    if (int e = planemap::make_plane_map(&a.gmap, dgate_planes, H, W, cy / 8, 1)) return e;
    a.H = H; a.W = W; a.inv_scale = inv_scale; a.dw = dweight;
    cudaStream_t st = (cudaStream_t)stream;
    if (cx == 48 && cy == 64) return launch_dw<6, 6, 8, 2>(a, st);     // cells 0, 3
    if (cx == 32 && cy == 64) return launch_dw<4, 4, 8, 3>(a, st);     // cells 1, 2
    if (cx == 40 && cy == 32) return launch_dw<5, 6, 4, 3>(a, st);     // cell 4
    return AARMVS_ERR_ARG;
}
