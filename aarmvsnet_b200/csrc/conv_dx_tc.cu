// Data gradient of the ConvLSTM cells on the tensor cores (BASELINE config 5; the backward of models/module.py:80-82):
//
//   dX[128 px][cin] = sum over 9 taps, K = 4*hidden:  dG_tap[128 px][4*hidden] * W'_tap[4*hidden][cin]
//
// the same TMA-fed tcgen05 / TMEM implicit GEMM as the forward cell (convlstm_tma.cu: operand planes, 16 x 8 pixel tiles, a
// staged 18 x 10 halo tile whose taps are start addresses, split-fp16 operands hi*hi + lo*hi + hi*lo, warp roles, mbarrier
// hand-offs), with the gate gradients as the A operand, the spatially flipped, transposed cell weights as B, and a plain
// epilogue that scatters the cin columns to the cell's inputs (stored or accumulated, as aarmvs_conv_nhwc does).
//
// Gradients are small numbers (1e-9 .. 1e-3) and fp16 underflows below 6e-5, so the planes hold 2^k * dG with k chosen per
// launch from max |dG| (aarmvs_split_planes_scaled: the maximum lands in [128, 256)), and the epilogue multiplies by 2^-k:
// the (hi, lo) pair then carries dG to 2^-22 relative for every element within 2^-15 of the largest, and to an absolute
// 2^-25 * 2^-k below that -- far inside the fp32 rounding of the sums it enters.
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

struct OutSeg {
    float* ptr;
    int ch, acc;
};

struct alignas(64) DxArgs {
    CUtensorMap map;            // gate-gradient operand planes (W*8 halfs, H, chunks, 2 = hi/lo, 1), box (80, 18, chunks, 1, 1)
    int H, W;
    const __half* weight;       // [9 * KSTEPS][2 chunks][2 (hi, lo)][NG][8] fp16
    const float* inv_scale;     // device: 2^-k of the planes' scale
    OutSeg out[3];
    int nout, cout;             // cout <= NG real output channels
};

template <int KSTEPS, int NG, int NST>
struct Cfg {
    static constexpr int KC8 = 2 * KSTEPS;
    static constexpr int NSTEPS = 9 * KSTEPS;
    static constexpr int A_PLANE = KC8 * CHUNK_BYTES;
    static constexpr int A_STAGE = 2 * A_PLANE;
    static constexpr int B_BYTES = NSTEPS * 2 * 2 * NG * 16;
    static constexpr int SMEM = B_BYTES + NST * A_STAGE + 1024;
    static constexpr int TMEM_COLS = 4 * NG <= 128 ? 128 : 256;              // two accumulators x (hi*hi + lo*hi | hi*lo)
    static_assert(4 * NG <= 256 && NG % 16 == 0, "accumulator columns");
    static_assert(B_BYTES % 128 == 0 && A_PLANE % 128 == 0, "TMA destinations must be 128-byte aligned");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// 32 lanes x 4 consecutive 32-bit columns, no wait (the caller waits once for all its loads)
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr)
                 : "memory");
}

template <int KSTEPS, int NG, int NST>
__global__ void __launch_bounds__(N_THREADS, 1) conv_dx_tc_kernel(const __grid_constant__ DxArgs a) {
    using C = Cfg<KSTEPS, NG, NST>;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;                                   // weights: [K step][2 chunks][hi rows | lo rows][16 B]
    unsigned char* sA = smem + C::B_BYTES;                      // stages: [stage][hi|lo][kc8][px][16 B]
    __shared__ uint64_t full_a[NST], empty_a[NST], tmem_full[2], tmem_empty[2], w_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
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
        tc::tma_prefetch_desc(&a.map);
        tc::mbar_arrive_expect_tx(&w_bar, C::B_BYTES);
        for (int off = 0; off < C::B_BYTES; off += 32768)
            tc::bulk_load_1d(sB + off, reinterpret_cast<const unsigned char*>(a.weight) + off, min(32768, C::B_BYTES - off), &w_bar);
    }
    const uint32_t tmem_base = tmem_base_s;
    const int tiles_x = (a.W + TCOL - 1) / TCOL, tiles_y = (a.H + TR - 1) / TR;
    const int ntiles = tiles_x * tiles_y;

    if (warp == TMA_WARP) {
        // =========================================== TMA producer ===========================================
        constexpr uint32_t bytes = 2u * CHUNK_BYTES * C::KC8;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int st = it % NST;
            if (lane == 0) {
                tc::mbar_wait(&empty_a[st], ((it / NST) & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&full_a[st], bytes);
            }
            __syncwarp();
            if (lane < 2) {                              // the hi and the lo plane of the tile: one warp instruction
                const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
                tc::tma_load_5d(sA + st * C::A_STAGE + lane * C::A_PLANE, &a.map, &full_a[st], (tx * TCOL - 1) * 8, ty * TR - 1, 0, lane, 0);
            }
            __syncwarp();
        }
    } else if (warp == MMA_WARP) {
        // =========================================== MMA issuer ===========================================
        const bool leader = tc::elect_one();
        const uint32_t idesc2 = tc::idesc_f16_f32(128, 2 * NG), idesc1 = tc::idesc_f16_f32(128, NG);
        const uint32_t a_base = tc::smem_u32(sA), b_base = tc::smem_u32(sB);
        constexpr uint32_t A_HI = tc::desc_hi(HC * 16), B_HI = tc::desc_hi(8 * 16);
        const uint32_t b_lo = tc::desc_lo(b_base, 2 * NG * 16);       // [W_hi | W_lo] stacked along N
        tc::mbar_wait(&w_bar, 0);
        int it = 0;
        if ((int)blockIdx.x < ntiles) {
            tc::mbar_wait(&full_a[0], 0);
            tc::mbar_wait(&tmem_empty[0], 1);
            tc::tc_fence_after();
        }
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int st = it % NST, ab = it & 1;
            const uint32_t dst = tmem_base + ab * 2 * NG;
            const uint32_t a_lo_hi = tc::desc_lo(a_base + st * C::A_STAGE, NPX * 16);
            const uint32_t a_lo_lo = tc::desc_lo(a_base + st * C::A_STAGE + C::A_PLANE, NPX * 16);
            const bool more = tile + (int)gridDim.x < ntiles;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks) {
                        const uint32_t ak = (uint32_t)(((2 * ks) * NPX + (tap / 3) * HC + (tap % 3)) * 16) >> 4;
                        const uint32_t bk = (uint32_t)((tap * KSTEPS + ks) * 2 * 2 * NG);
                        tc::mma_f16_parts(dst, a_lo_hi + ak, A_HI, b_lo + bk, B_HI, idesc2, (tap | ks) != 0);   // hi*hi | hi*lo
                        tc::mma_f16_parts(dst, a_lo_lo + ak, A_HI, b_lo + bk, B_HI, idesc1, 1);                 // lo*hi
                    }
                }
                if (tap == 5 && more) {          // next tile's operands and accumulator
                    tc::mbar_wait(&full_a[(it + 1) % NST], ((it + 1) / NST) & 1);
                    tc::mbar_wait(&tmem_empty[(it + 1) & 1], (((it + 1) >> 1) & 1) ^ 1);
                    tc::tc_fence_after();
                }
            }
            if (leader) {
                tc::mma_commit(&empty_a[st]);
                tc::mma_commit(&tmem_full[ab]);
            }
            __syncwarp();
        }
    } else {
        // =========================================== epilogue ===========================================
        const int q = warp & 3, part = warp >> 2;                // TMEM lane quadrant, quarter of the output columns
        constexpr int NCOL = NG / 4;                             // 12 or 8 columns per thread
        const int m = q * 32 + lane;
        const float inv = *a.inv_scale;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
            const int y = ty * TR + (m >> 3), x = tx * TCOL + (m & 7);
            const bool valid = y < a.H && x < a.W;
            const size_t pix = valid ? (size_t)y * a.W + x : 0;
            tc::mbar_wait_relaxed(&tmem_full[ab], (it >> 1) & 1);
            tc::tc_fence_after();
            uint32_t g[NCOL / 4][4], g2[NCOL / 4][4];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 2 * NG + part * NCOL;
#pragma unroll
            for (int j = 0; j < NCOL / 4; ++j) {
                tmem_ld4_nowait(taddr + j * 4, g[j]);
                tmem_ld4_nowait(taddr + NG + j * 4, g2[j]);
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc::tc_fence_before();
            tc::mbar_arrive(&tmem_empty[ab]);                    // accumulator is in registers: release it
            if (!valid) continue;
#pragma unroll
            for (int j = 0; j < NCOL / 4; ++j) {
                const int co = part * NCOL + j * 4;
                if (co >= a.cout) continue;
                int base = 0, sg = 0;
                while (sg < a.nout - 1 && co >= base + a.out[sg].ch) { base += a.out[sg].ch; ++sg; }
                float* p = a.out[sg].ptr + pix * a.out[sg].ch + (co - base);
                float4 o = make_float4((__uint_as_float(g[j][0]) + __uint_as_float(g2[j][0])) * inv,
                                       (__uint_as_float(g[j][1]) + __uint_as_float(g2[j][1])) * inv,
                                       (__uint_as_float(g[j][2]) + __uint_as_float(g2[j][2])) * inv,
                                       (__uint_as_float(g[j][3]) + __uint_as_float(g2[j][3])) * inv);
                if (a.out[sg].acc) {
                    const float4 t = ld4(p);
                    o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
                }
                st4(p, o);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc<C::TMEM_COLS>(tmem_base);
}

template <int KSTEPS, int NG, int NST>
int launch_dx(const DxArgs& a, cudaStream_t stream) {
    using C = Cfg<KSTEPS, NG, NST>;
    auto kern = conv_dx_tc_kernel<KSTEPS, NG, NST>;
    static unsigned long long smem_ok = 0;
    if (cudaError_t e = ensure_dyn_smem(kern, C::SMEM, smem_ok)) return (int)e;
    const int ntiles = ((a.W + TCOL - 1) / TCOL) * ((a.H + TR - 1) / TR);
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    kern<<<balanced_grid(ntiles, sms), N_THREADS, C::SMEM, stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}

// fp32 [P][C] -> operand planes [2][C/8][P][8] of 2^k * x, k from the maximum the producer left in *max_bits (the bit pattern
// of max |x|, a non-negative float: unsigned order = float order): the largest element lands in [128, 256).
__global__ void __launch_bounds__(256) split_planes_scaled_kernel(const float* __restrict__ x, size_t P, int C,
                                                                  const unsigned* __restrict__ max_bits, __half* __restrict__ planes,
                                                                  float* __restrict__ inv_scale) {
    const float mx = __uint_as_float(*max_bits);
    int e = 0;
    if (mx > 0.f && mx < 3.0e38f) frexpf(mx, &e);               // mx = f * 2^e, f in [0.5, 1)
    const int k = mx > 0.f ? max(-100, min(100, 8 - e)) : 0;
    const float s = ldexpf(1.f, k);
    if (blockIdx.x == 0 && threadIdx.x == 0) *inv_scale = ldexpf(1.f, -k);
    const int chunks = C / 8;
    const size_t n = P * chunks;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t p = i / chunks;
        const int c = (int)(i - p * chunks);
        const float* xp = x + p * C + c * 8;
        float4 v0 = ld4(xp), v1 = ld4(xp + 4);
        v0.x *= s; v0.y *= s; v0.z *= s; v0.w *= s;
        v1.x *= s; v1.y *= s; v1.z *= s; v1.w *= s;
        uint2 h0, l0, h1, l1;
        tc::split4_f16(v0, h0, l0);
        tc::split4_f16(v1, h1, l1);
        *reinterpret_cast<uint4*>(planes + ((size_t)c * P + p) * 8) = make_uint4(h0.x, h0.y, h1.x, h1.y);
        *reinterpret_cast<uint4*>(planes + ((size_t)(chunks + c) * P + p) * 8) = make_uint4(l0.x, l0.y, l1.x, l1.y);
    }
}

}  // namespace

extern "C" int aarmvs_split_planes_scaled(const float* x, int H, int W, int channels, const unsigned* max_bits, void* planes,
                                          float* inv_scale, void* stream) {
    AARMVS_CHECK_ARG(x && max_bits && planes && inv_scale && H > 0 && W > 0 && channels > 0 && channels % 8 == 0);
    AARMVS_CHECK_ARG(((uintptr_t)planes & 15) == 0 && ((uintptr_t)x & 15) == 0);
    const size_t P = (size_t)H * W;
    const size_t blocks = (P * (channels / 8) + 255) / 256;
    split_planes_scaled_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(x, P, channels, max_bits,
                                                                                                    (__half*)planes, inv_scale);
    AARMVS_LAUNCH_RESULT();
}

extern "C" int aarmvs_conv_dx_tc(const void* dgate_planes, int H, int W, int cin, const aarmvs_seg* out, int nout, int cout,
                                 const void* weight_f16, const float* inv_scale, void* stream) {
    AARMVS_CHECK_ARG(dgate_planes && out && weight_f16 && inv_scale && H > 0 && W > 0 && nout >= 1 && nout <= 3);
    AARMVS_CHECK_ARG(((uintptr_t)dgate_planes & 15) == 0 && ((uintptr_t)weight_f16 & 15) == 0);
    DxArgs a = {};
    int csum = 0;
    for (int s = 0; s < nout; ++s) {
        AARMVS_CHECK_ARG(out[s].ptr && out[s].channels > 0 && out[s].channels % 4 == 0 && ((uintptr_t)out[s].ptr & 15) == 0);
        a.out[s].ptr = out[s].ptr; a.out[s].ch = out[s].channels; a.out[s].acc = out[s].accumulate;
        csum += out[s].channels;
    }
    AARMVS_CHECK_ARG(csum == cout);
    if (int e = planemap::make_plane_map(&a.map, dgate_planes, H, W, cin / 8, 1)) return e;
    a.H = H; a.W = W; a.weight = (const __half*)weight_f16; a.inv_scale = inv_scale; a.nout = nout; a.cout = cout;
    cudaStream_t st = (cudaStream_t)stream;
    if (cin == 64 && cout == 48) return launch_dx<4, 48, 2>(a, st);     // cells 0, 3
    if (cin == 64 && cout == 32) return launch_dx<4, 32, 3>(a, st);     // cells 1, 2
    if (cin == 32 && cout == 40) return launch_dx<2, 48, 4>(a, st);     // cell 4 (output columns padded 40 -> 48 with zero weights)
    return AARMVS_ERR_ARG;
}
