// deConvGnReLU (models/module.py:269-287) on the tensor cores: ConvTranspose2d(16, 16, k3, s2, p1, op1) as a TMA-fed
// tcgen05 implicit GEMM over the fp16 (hi, lo) operand planes of h (written by the ConvLSTM epilogue), in two passes
// because GroupNorm needs the sums of the WHOLE output map:
//   pass 1 (out_planes == NULL): transposed conv -> GroupNorm sums only (nothing is stored);
//   pass 2: the same GEMM again -> GroupNorm + ReLU -> fp16 split -> operand planes of the next cell.
// Recomputing 2.2 GFLOP on the tensor pipe costs less than writing and re-reading the fp32 map (the CUDA-core
// deconvolution + aarmvs_split_planes pair this replaces took 0.16 ms per plane at cfg2, FMA pipe 28 % busy).
//
// out[2i+a][2j+b] gathers the input pixels (i,j), (i,j+1), (i+1,j), (i+1,j+1)   (from y = 2 i - 1 + ky):
//   class (0,0): in(i,j) W[1][1]
//   class (0,1): in(i,j) W[1][2] + in(i,j+1) W[1][0]
//   class (1,0): in(i,j) W[2][1] + in(i+1,j) W[0][1]
//   class (1,1): in(i,j) W[2][2] + in(i,j+1) W[2][0] + in(i+1,j) W[0][2] + in(i+1,j+1) W[0][0]
// i.e. 9 (class, tap) pairs, each one K = 16 step over the 16 input channels.  M = the 128 input pixels of a 16 x 8 tile,
// the staged box is [18][10] pixels starting AT the tile origin (rows / columns +1 are the needed neighbours; TMA's zero
// fill beyond the image is the transposed conv's implicit zero), and a tap is a descriptor start address, exactly as in
// convlstm_tma.cu.  Per pair: A_hi x [W_hi | W_lo] (N = 32) and A_lo x W_hi (N = 16) into the class's 32 TMEM columns.
#include "common.cuh"
#include "tc_common.cuh"
#include "plane_map.cuh"
#include "../../include/aarmvs.h"

namespace {

constexpr int TR = 16, TCOL = 8, HR = planemap::BOX_ROWS, HC = planemap::BOX_COLS, NPX = HR * HC;
constexpr int CHUNK_BYTES = NPX * 16;
constexpr int A_PLANE = 2 * CHUNK_BYTES, A_STAGE = 2 * A_PLANE;          // 16 channels = 2 chunks; hi and lo planes
constexpr int NST = 4;
constexpr int NPAIR = 9, PAIR_BYTES = 2 * 32 * 16;                         // [2 chunks][16 W_hi rows | 16 W_lo rows][16 B]
constexpr int B_BYTES = NPAIR * PAIR_BYTES;                                // 9216
constexpr int SMEM = B_BYTES + NST * A_STAGE + 1024;
constexpr int N_EPI_WARPS = 8, MMA_WARP = N_EPI_WARPS, TMA_WARP = MMA_WARP + 1;
constexpr int N_THREADS = (N_EPI_WARPS + 2) * 32;
constexpr int TMEM_COLS = 256;                                             // 2 accumulators x 4 classes x 32 columns
static_assert(B_BYTES % 128 == 0 && A_PLANE % 128 == 0 && A_STAGE % 128 == 0, "TMA destinations must be 128-byte aligned");

// (class, tap) pairs in issue order: class = a * 2 + b, (dy, dx) = offset of the input pixel
__device__ constexpr int PAIR_CLS[NPAIR] = {0, 1, 1, 2, 2, 3, 3, 3, 3};
__device__ constexpr int PAIR_OFF[NPAIR] = {0, 0, 1, 0, HC, 0, 1, HC, HC + 1};       // dy * HC + dx, in pixels
__device__ constexpr int PAIR_FIRST[NPAIR] = {1, 1, 0, 1, 0, 1, 0, 0, 0};

struct alignas(64) DeconvTcArgs {
    CUtensorMap map;            // input planes: (W*8 halfs, H, 2 chunks, 2 = hi/lo, ping-pong halves)
    int pp_phase;               // < 0: single buffer; else the half to read is (d + pp_phase) & 1
    int Hin, Win;
    const __half* weight;       // [9 pairs][2 chunks][32 rows: 0-15 hi(W[.][co]), 16-31 lo][8] fp16
    float bias[16];
    double* stats;              // [2 groups][sum, sumsq]: accumulated by pass 1, read by pass 2
    const float* gamma;
    const float* beta;
    double count;               // elements per GroupNorm group (8 channels x output pixels)
    __half* out_planes;         // pass 2: [2][2 chunks][2 Hin][2 Win][8]; NULL = pass 1
    int d_host;
    const int* d_dev;
};

__global__ void __launch_bounds__(N_THREADS, 1) deconv_tc_kernel(const __grid_constant__ DeconvTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;
    unsigned char* sA = smem + B_BYTES;
    __shared__ uint64_t full_a[NST], empty_a[NST], tmem_full[2], tmem_empty[2], w_bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_scale[16], s_shift[16];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool apply = a.out_planes != nullptr;

    // setup that reads nothing a predecessor writes (overlaps the previous kernel's drain under programmatic dependent launch)
    if (tid == 0) {
        for (int s = 0; s < NST; ++s) { tc::mbar_init(&full_a[s], 1); tc::mbar_init(&empty_a[s], 1); }
        for (int b = 0; b < 2; ++b) { tc::mbar_init(&tmem_full[b], 1); tc::mbar_init(&tmem_empty[b], N_EPI_WARPS * 32); }
        tc::mbar_init(&w_bar, 1);
        tc::mbar_fence_init();
    }
    if (warp == MMA_WARP) tc::tmem_alloc<TMEM_COLS>(&tmem_base_s);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (warp == TMA_WARP && lane == 0) {
        tc::tma_prefetch_desc(&a.map);
        tc::mbar_arrive_expect_tx(&w_bar, B_BYTES);
        tc::bulk_load_1d(sB, a.weight, B_BYTES, &w_bar);
    }
    tc::pdl_wait();
    tc::pdl_launch_dependents();
    const int d = plane_index(a.d_host, a.d_dev);
    if (tid < 16) {          // GroupNorm scale / shift from the sums pass 1 (the predecessor) accumulated
        float sc = 1.f, sh = 0.f;
        if (apply) {
            float mean, rstd;
            gn_mean_rstd(a.stats + (tid >> 3) * 2, a.count, mean, rstd);
            sc = rstd * a.gamma[tid];
            sh = a.beta[tid] - mean * sc;
        }
        s_scale[tid] = sc;
        s_shift[tid] = sh;
    }
    __syncthreads();
    const uint32_t tmem_base = tmem_base_s;

    const int tiles_x = (a.Win + TCOL - 1) / TCOL, tiles_y = (a.Hin + TR - 1) / TR;
    const int ntiles = tiles_x * tiles_y;

    if (warp == TMA_WARP) {
        const int sel = a.pp_phase < 0 ? 0 : ((d + a.pp_phase) & 1);
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int st = it % NST;
            if (lane == 0) {
                tc::mbar_wait(&empty_a[st], ((it / NST) & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&full_a[st], A_STAGE);
            }
            __syncwarp();
            if (lane < 2) {          // hi and lo planes: both copies in one warp instruction
                const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
                tc::tma_load_5d(sA + st * A_STAGE + lane * A_PLANE, &a.map, &full_a[st], tx * TCOL * 8, ty * TR, 0, lane, sel);
            }
            __syncwarp();
        }
    } else if (warp == MMA_WARP) {
        const bool leader = tc::elect_one();
        const uint32_t idesc2 = tc::idesc_f16_f32(128, 32), idesc1 = tc::idesc_f16_f32(128, 16);
        const uint32_t a_base = tc::smem_u32(sA), b_base = tc::smem_u32(sB);
        constexpr uint32_t A_HI = tc::desc_hi(HC * 16), B_HI = tc::desc_hi(8 * 16);
        const uint32_t b_lo = tc::desc_lo(b_base, 32 * 16);
        tc::mbar_wait(&w_bar, 0);
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int st = it % NST, ab = it & 1;
            tc::mbar_wait(&full_a[st], (it / NST) & 1);
            tc::mbar_wait(&tmem_empty[ab], ((it >> 1) & 1) ^ 1);
            tc::tc_fence_after();
            const uint32_t a_lo_hi = tc::desc_lo(a_base + st * A_STAGE, NPX * 16);
            const uint32_t a_lo_lo = tc::desc_lo(a_base + st * A_STAGE + A_PLANE, NPX * 16);
            if (leader) {
#pragma unroll
                for (int p = 0; p < NPAIR; ++p) {
                    const uint32_t dst = tmem_base + ab * 128 + PAIR_CLS[p] * 32;
                    const uint32_t ak = (uint32_t)PAIR_OFF[p], bk = (uint32_t)(p * (PAIR_BYTES / 16));
                    tc::mma_f16_parts(dst, a_lo_hi + ak, A_HI, b_lo + bk, B_HI, idesc2, !PAIR_FIRST[p]);
                    tc::mma_f16_parts(dst, a_lo_lo + ak, A_HI, b_lo + bk, B_HI, idesc1, 1);
                }
                tc::mma_commit(&empty_a[st]);
                tc::mma_commit(&tmem_full[ab]);
            }
            __syncwarp();
        }
    } else {
        // epilogue: thread = input pixel m of the tile (TMEM lane), warps 0-3 take output rows 2i (classes 0, 1), warps 4-7
        // rows 2i+1 (classes 2, 3)
        const int q = warp & 3, arow = warp >> 2;
        const int m = q * 32 + lane;
        const int Wo = 2 * a.Win;
        const size_t Po = (size_t)4 * a.Hin * a.Win;
        float gs[2] = {0.f, 0.f}, gq[2] = {0.f, 0.f};
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
            const int i = ty * TR + (m >> 3), j = tx * TCOL + (m & 7);
            const bool valid = i < a.Hin && j < a.Win;
            tc::mbar_wait_relaxed(&tmem_full[ab], (it >> 1) & 1);
            tc::tc_fence_after();
            float acc[2][32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + ab * 128 + arow * 64;
            tc::tmem_ld32(taddr, acc[0]);
            tc::tmem_ld32(taddr + 32, acc[1]);
            tc::tc_fence_before();
            tc::mbar_arrive(&tmem_empty[ab]);
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                float v[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = (acc[b][c] + acc[b][16 + c]) + a.bias[c];      // (hi*hi + lo*hi) + hi*lo, + bias
                if (!apply) {
                    if (valid) {
#pragma unroll
                        for (int c = 0; c < 16; c += 4) {
                            gs[c >> 3] += (v[c] + v[c + 1]) + (v[c + 2] + v[c + 3]);
                            gq[c >> 3] += (v[c] * v[c] + v[c + 1] * v[c + 1]) + (v[c + 2] * v[c + 2] + v[c + 3] * v[c + 3]);
                        }
                    }
                } else if (valid) {
                    const size_t opix = (size_t)(2 * i + arow) * Wo + 2 * j + b;
#pragma unroll
                    for (int ch = 0; ch < 2; ++ch) {
                        float y[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) y[c] = fmaxf(fmaf(v[ch * 8 + c], s_scale[ch * 8 + c], s_shift[ch * 8 + c]), 0.f);
                        uint2 h0, l0, h1, l1;
                        tc::split4_f16(make_float4(y[0], y[1], y[2], y[3]), h0, l0);
                        tc::split4_f16(make_float4(y[4], y[5], y[6], y[7]), h1, l1);
                        *reinterpret_cast<uint4*>(a.out_planes + ((size_t)ch * Po + opix) * 8) = make_uint4(h0.x, h0.y, h1.x, h1.y);
                        *reinterpret_cast<uint4*>(a.out_planes + ((size_t)(2 + ch) * Po + opix) * 8) = make_uint4(l0.x, l0.y, l1.x, l1.y);
                    }
                }
            }
        }
        if (!apply) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const double s = warp_sum((double)gs[g]), qq = warp_sum((double)gq[g]);
                if (lane == 0) {
                    atomicAdd(a.stats + g * 2 + 0, s);
                    atomicAdd(a.stats + g * 2 + 1, qq);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace

extern "C" int aarmvs_deconv2x_tc(const aarmvs_plane_src* in, int Hin, int Win, const void* weight_f16, const float* bias_host,
                                  double* gn_sums, const float* gn_weight, const float* gn_bias, void* out_planes, int d_host,
                                  const int* d_dev, void* stream) {
    AARMVS_CHECK_ARG(in && in->planes && in->channels == 16 && (in->pp == 1 || in->pp == 2) && Hin > 0 && Win > 0);
    AARMVS_CHECK_ARG(weight_f16 && bias_host && gn_sums && ((uintptr_t)in->planes & 15) == 0 && ((uintptr_t)weight_f16 & 15) == 0);
    AARMVS_CHECK_ARG(!out_planes || (gn_weight && gn_bias && ((uintptr_t)out_planes & 15) == 0));
    DeconvTcArgs a = {};
    if (int e = planemap::make_plane_map(&a.map, in->planes, Hin, Win, 2, in->pp)) return e;
    a.pp_phase = in->pp == 2 ? (in->pp_phase & 1) : -1;
    a.Hin = Hin; a.Win = Win; a.weight = (const __half*)weight_f16;
    for (int i = 0; i < 16; ++i) a.bias[i] = bias_host[i];
    a.stats = gn_sums; a.gamma = gn_weight; a.beta = gn_bias; a.count = 8.0 * 4.0 * (double)Hin * (double)Win;
    a.out_planes = (__half*)out_planes; a.d_host = d_host; a.d_dev = d_dev;
    static unsigned long long smem_ok = 0;
    if (cudaError_t e = ensure_dyn_smem(deconv_tc_kernel, SMEM, smem_ok)) return (int)e;
    const int ntiles = ((Win + TCOL - 1) / TCOL) * ((Hin + TR - 1) / TR);
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    if (cudaError_t e = launch_pdl(deconv_tc_kernel, dim3(balanced_grid(ntiles, sms)), dim3(N_THREADS), SMEM, (cudaStream_t)stream, a))
        return (int)e;
    AARMVS_LAUNCH_RESULT();
}
