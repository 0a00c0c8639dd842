// Cost head on the tensor cores: conv_0 (3x3, 8 -> 1; models/drmvsnet.py:165) as a TMA-fed tcgen05 implicit GEMM over
// the fp16 (hi, lo) operand planes of h4 (written by cell 4's epilogue), fused with the winner-take-all update
// (drmvsnet.py:324-339):  p = exp(cost); if (max_prob < p) {max_prob = p; depth = depth_values[d]; index = d};
// exp_sum += p; conf = max_prob / exp_sum at the last plane.
//
// h4 has 8 channels = ONE 16-byte K chunk per pixel, the K = 16 instruction wants two: an MMA takes TWO TAPS, the second
// chunk being the leading-byte-offset away from the first ((0,0)+(0,1), (0,2)+(1,0), (1,1)+(1,2), (2,0)+(2,1), (2,2)+itself
// with zero weights).  B rows 0 / 8 hold hi / lo of the weights, so with N = 16 the sum of accumulator columns 0 and 8
// is the full (hi + lo)(hi + lo) product: 5 pairs x (A_hi, A_lo) = 10 MMAs per 128-pixel tile.
#include "common.cuh"
#include "tc_common.cuh"
#include "plane_map.cuh"
#include "../../include/aarmvs.h"

namespace {

constexpr int TR = 16, TCOL = 8, HR = planemap::BOX_ROWS, HC = planemap::BOX_COLS, NPX = HR * HC;
constexpr int A_PLANE = (NPX * 16 + 127) / 128 * 128, A_STAGE = 2 * A_PLANE;   // one chunk (padded to 128 B); hi and lo planes
constexpr int NST = 4, NPAIR = 5, PAIR_BYTES = 2 * 16 * 16;                // [2 taps][16 rows][16 B]
constexpr int B_BYTES = NPAIR * PAIR_BYTES;
constexpr int SMEM = 4096 + NST * A_STAGE + 1024;                          // weights padded to 4 KB: stages stay 128-byte aligned
constexpr int N_EPI_WARPS = 4, MMA_WARP = N_EPI_WARPS, TMA_WARP = MMA_WARP + 1;
constexpr int N_THREADS = (N_EPI_WARPS + 2) * 32;
constexpr int TMEM_COLS = 64;                                              // 2 accumulators, 32 columns apart
constexpr int CTAS_PER_SM = 4;            // the per-tile chain (state load -> accumulator -> exp -> stores) is latency bound:
                                          // four small CTAs per SM (30 KB smem, 64 TMEM columns each) keep four tiles in flight
static_assert(B_BYTES <= 4096 && A_PLANE % 128 == 0, "TMA destinations must be 128-byte aligned");

__device__ constexpr int PAIR_A[NPAIR] = {0, 2, HC + 1, 2 * HC, 2 * HC + 2};              // first tap: ky * HC + kx
__device__ constexpr int PAIR_LBO[NPAIR] = {1, HC - 2, 1, 1, 0};                          // second tap - first tap, in pixels

struct alignas(64) HeadTcArgs {
    CUtensorMap map;
    int H, W;
    const __half* weight;       // [5 pairs][2 taps][16 rows: row 0 hi(w), row 8 lo(w)][8] fp16
    float bias;
    const float* depth_values;
    int num_depth, d_host;
    const int* d_dev;
    float *max_prob, *depth_map, *exp_sum, *conf, *cost_volume;
    int* index_map;
    // optional fused training loss (mvsnet_cls_loss, models/drmvsnet.py:347-381, without the [D][H][W] probability volume)
    const int* gt_index;        // [H*W] hypothesis nearest to the ground truth (0 where the mask is 0)
    const float* mask;          // [H*W]
    float* gt_cost;             // [3][H*W] state: regularised cost at the ground-truth plane | running max | rescaled exp-sum
    double* loss_sum;           // += sum over pixels of mask * (log(sum_d exp(cost_d)) - cost_gt) at the last plane
};

__global__ void __launch_bounds__(N_THREADS, CTAS_PER_SM) head_tc_kernel(const __grid_constant__ HeadTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;
    unsigned char* sA = smem + 4096;
    __shared__ uint64_t full_a[NST], empty_a[NST], tmem_full[2], tmem_empty[2], w_bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // setup that reads nothing a predecessor writes (overlaps cell 4's drain under programmatic dependent launch)
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
    const uint32_t tmem_base = tmem_base_s;
    const int tiles_x = (a.W + TCOL - 1) / TCOL, tiles_y = (a.H + TR - 1) / TR;
    const int ntiles = tiles_x * tiles_y;

    if (warp == TMA_WARP) {
        const int sel = d & 1;                                   // h4 of the CURRENT plane
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int st = it % NST;
            if (lane == 0) {
                tc::mbar_wait(&empty_a[st], ((it / NST) & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&full_a[st], 2 * NPX * 16);
            }
            __syncwarp();
            if (lane < 2) {          // hi and lo planes: both copies in one warp instruction
                const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
                tc::tma_load_5d(sA + st * A_STAGE + lane * A_PLANE, &a.map, &full_a[st], (tx * TCOL - 1) * 8, ty * TR - 1, 0, lane, sel);
            }
            __syncwarp();
        }
    } else if (warp == MMA_WARP) {
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::idesc_f16_f32(128, 16);
        const uint32_t a_base = tc::smem_u32(sA), b_base = tc::smem_u32(sB);
        constexpr uint32_t A_HI = tc::desc_hi(HC * 16), B_HI = tc::desc_hi(8 * 16);
        const uint32_t b_lo = tc::desc_lo(b_base, 16 * 16);
        tc::mbar_wait(&w_bar, 0);
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int st = it % NST, ab = it & 1;
            tc::mbar_wait(&full_a[st], (it / NST) & 1);
            tc::mbar_wait(&tmem_empty[ab], ((it >> 1) & 1) ^ 1);
            tc::tc_fence_after();
            if (leader) {
                const uint32_t dst = tmem_base + ab * 32;
                const uint32_t hi_addr = a_base + st * A_STAGE, lo_addr = hi_addr + A_PLANE;
#pragma unroll
                for (int p = 0; p < NPAIR; ++p) {
                    const uint32_t bk = (uint32_t)(p * (PAIR_BYTES / 16));
                    tc::mma_f16_parts(dst, tc::desc_lo(hi_addr + PAIR_A[p] * 16, PAIR_LBO[p] * 16), A_HI, b_lo + bk, B_HI, idesc, p != 0);
                    tc::mma_f16_parts(dst, tc::desc_lo(lo_addr + PAIR_A[p] * 16, PAIR_LBO[p] * 16), A_HI, b_lo + bk, B_HI, idesc, 1);
                }
                tc::mma_commit(&empty_a[st]);
                tc::mma_commit(&tmem_full[ab]);
            }
            __syncwarp();
        }
    } else {
        const int m = warp * 32 + lane;
        const float depth = a.depth_values[d];
        const size_t n = (size_t)a.H * a.W;
        float ce_sum = 0.f;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int ab = it & 1;
            const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
            const int y = ty * TR + (m >> 3), x = tx * TCOL + (m & 7);
            const bool valid = y < a.H && x < a.W;
            const size_t p = valid ? (size_t)y * a.W + x : 0;
            // the WTA state of this pixel: requested before waiting for the accumulator
            const float mp0 = (d == 0 || !valid) ? 0.f : a.max_prob[p];
            const float es0 = (d == 0 || !valid) ? 0.f : a.exp_sum[p];
            tc::mbar_wait_relaxed(&tmem_full[ab], (it >> 1) & 1);
            tc::tc_fence_after();
            uint32_t r[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + ab * 32;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc::tc_fence_before();
            tc::mbar_arrive(&tmem_empty[ab]);
            if (!valid) continue;
            const float acc = (__uint_as_float(r[0]) + __uint_as_float(r[8])) + a.bias;
            float gc = 0.f, lse = 0.f;
            if (a.gt_index) {
                const int gi = a.gt_index[p];
                if (d == gi) { gc = acc; a.gt_cost[p] = acc; }
                else if (d == a.num_depth - 1) gc = a.gt_cost[p];
                // log-sum-exp of the costs as an online softmax (running max + rescaled sum), like the stable F.softmax of the
                // reference's loss path (drmvsnet.py:291): the winner-take-all exp-sum below may overflow, this cannot
                const float m0 = d == 0 ? -INFINITY : a.gt_cost[n + p];
                const float s0 = d == 0 ? 0.f : a.gt_cost[2 * n + p];
                const float m1 = fmaxf(m0, acc);
                const float s1 = s0 * expf(m0 - m1) + expf(acc - m1);
                a.gt_cost[n + p] = m1;
                a.gt_cost[2 * n + p] = s1;
                lse = m1 + logf(s1);
            }
            if (a.cost_volume) a.cost_volume[(size_t)d * n + p] = acc;
            const float prob = expf(acc);
            float mp = mp0;
            if (d == 0) { a.depth_map[p] = 0.f; a.index_map[p] = 0; }
            if (mp < prob) {           // strict '<': the first plane wins ties (drmvsnet.py:327)
                mp = prob;
                a.depth_map[p] = depth;
                a.index_map[p] = d;
            }
            const float es = __fadd_rn(es0, prob);
            a.max_prob[p] = mp;
            a.exp_sum[p] = es;
            if (d == a.num_depth - 1) {
                a.conf[p] = __fdiv_rn(mp, es);
                // -log softmax(cost)[gt] = log(sum_d exp(cost_d)) - cost_gt, times the mask (drmvsnet.py:366-369)
                if (a.gt_index) ce_sum += a.mask[p] * (lse - gc);
            }
        }
        if (a.gt_index && d == a.num_depth - 1) {
            const double s = warp_sum((double)ce_sum);
            if (lane == 0) atomicAdd(a.loss_sum, s);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tc::tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace

extern "C" int aarmvs_head_wta_tc(const void* h4_planes, int H, int W, const void* weight_f16, const float* bias_host,
                                  const float* depth_values, int num_depth, int d_host, const int* d_dev, float* max_prob,
                                  float* depth_map, float* exp_sum, int* index_map, float* conf, float* cost_volume,
                                  const int* gt_index, const float* gt_mask, float* gt_cost, double* loss_sum, void* stream) {
    AARMVS_CHECK_ARG(h4_planes && weight_f16 && bias_host && depth_values && max_prob && depth_map && exp_sum && index_map && conf);
    AARMVS_CHECK_ARG(H > 0 && W > 0 && num_depth > 0 && ((uintptr_t)h4_planes & 15) == 0 && ((uintptr_t)weight_f16 & 15) == 0);
    HeadTcArgs a = {};
    if (int e = planemap::make_plane_map(&a.map, h4_planes, H, W, 1, 2)) return e;
    a.H = H; a.W = W; a.weight = (const __half*)weight_f16; a.bias = bias_host[0];
    a.depth_values = depth_values; a.num_depth = num_depth; a.d_host = d_host; a.d_dev = d_dev;
    a.max_prob = max_prob; a.depth_map = depth_map; a.exp_sum = exp_sum; a.index_map = index_map; a.conf = conf;
    a.cost_volume = cost_volume;
    AARMVS_CHECK_ARG(!gt_index || (gt_mask && gt_cost && loss_sum));
    a.gt_index = gt_index; a.mask = gt_mask; a.gt_cost = gt_cost; a.loss_sum = loss_sum;
    static unsigned long long smem_ok = 0;
    if (cudaError_t e = ensure_dyn_smem(head_tc_kernel, SMEM, smem_ok)) return (int)e;
    const int ntiles = ((W + TCOL - 1) / TCOL) * ((H + TR - 1) / TR);
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    const int grid = ntiles < sms * CTAS_PER_SM ? ntiles : sms * CTAS_PER_SM;
    if (cudaError_t e = launch_pdl(head_tc_kernel, dim3(grid), dim3(N_THREADS), SMEM, (cudaStream_t)stream, a)) return (int)e;
    AARMVS_LAUNCH_RESULT();
}
