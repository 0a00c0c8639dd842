// Modulated deformable 3x3 conv 32 -> 32 (models/module.py:133-189, DeformConv2d.forward) on the tensor cores.
//
//   out[p][o] = b[o] + sum_{tap k, channel c} v_k[p][c] * w[o][c][k],   v_k[p][c] = m_k(p) * bilinear_k(x)(p)[c]
//
// i.e. a GEMM with M = pixels, K = 9 taps x 32 channels, N = 32, whose A operand is GATHERED: per pixel and tap four
// clamped corner lines (module.py:150-165).  The structure is pass A's (cost_stem_tc.cu): a ring of independent worker
// groups, each owning an operand stage and a TMEM accumulator.  A group takes a tile of 16 x 8 pixels (one thread per
// pixel holds its 27 offsets / modulation logits; a 2-D tile keeps the 180 source lines its nine taps share in L1 -- with
// 128 consecutive pixels of one row the footprint was 390 lines and the L1 hit rate 36 %) and walks the nine taps: write the gather recipes of the
// tap, gather with 8 lanes per pixel (one fully used 128-byte line per corner -- the CUDA-core kernel this replaces had
// one THREAD per pixel: 32 L1 wavefronts per load instruction, 3.5 ms per full-resolution layer), apply the input's
// GroupNorm + ReLU, blend in the reference's operation order, split to fp16 (hi, lo), stage a K-major operand; the MMA
// warp issues A_hi x [W_hi | W_lo] (N = 64) and A_lo x W_hi (N = 32) per 16 channels into the group's accumulator; after
// tap 8 the group reads it, adds the bias, stores the pre-norm output and accumulates its GroupNorm sums.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/aarmvs.h"

namespace {

constexpr int G = 5, GROUP_THREADS = 128, N_THREADS = 32 + G * GROUP_THREADS;
constexpr int TILE = 128, TILE_R = 16, TILE_C = 8;                     // pixels per tile = accumulator rows; row m = pixel (m / 8, m % 8)
constexpr int AROWS = 130;                                             // operand rows per K chunk (= 2 mod 8: conflict-free stores)
constexpr int KC8 = 4;                                                 // 32 channels of one tap
constexpr int A_PLANE = KC8 * AROWS * 16, A_STAGE = 2 * A_PLANE;       // 8320, 16640
constexpr int STEP_BYTES = 2 * 64 * 16;                                // one K = 16 step of B: [2 chunks][32 hi rows | 32 lo rows][16 B]
constexpr int B_BYTES = 9 * 2 * STEP_BYTES;                            // 36864
constexpr int REC_BYTES = TILE * 48;
constexpr int SMEM = B_BYTES + G * (A_STAGE + REC_BYTES) + 1024;
constexpr int TMEM_COLS = 512;                                         // G accumulators x 64 columns (power of two)
static_assert(G * 64 <= TMEM_COLS, "accumulators exceed TMEM");

struct DeformTcArgs {
    const float* x;            // layer input NHWC [H][W][32], pre-norm (GroupNorm + ReLU applied on load) or plain
    const double* x_stats;
    const float* x_gamma;
    const float* x_beta;
    double x_count;
    int x_groups, x_mode;
    int H, W;
    const float* offmask;      // [H][W][32]: 0-8 row offsets, 9-17 column offsets, 18-26 modulation logits
    const __half* weight;      // [9 taps][2 K steps][2 chunks][64 rows: 0-31 hi(w[o]), 32-63 lo][8] fp16
    const float* bias;
    float* out;
    double* out_stats;         // 4 groups
};

// Gather recipe of one (pixel, tap): byte offsets of the four corner lines in the order lt, rb, lb, rt (0 with weight 0 for
// corners on the zero padding ring), their weights and the modulation.
struct __align__(16) Rec {
    unsigned o[4];
    float g[4];
    float m, pad0, pad1, pad2;
};

__global__ void __launch_bounds__(N_THREADS, 1) deform_tc_kernel(const __grid_constant__ DeformTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sB = smem;
    unsigned char* sA = smem + B_BYTES;
    unsigned char* sR = sA + G * A_STAGE;
    __shared__ uint64_t full_a[G], done_a[G];
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_scale[32], s_shift[32], s_bias[32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < B_BYTES / 16; i += N_THREADS)
        reinterpret_cast<uint4*>(sB)[i] = __ldg(reinterpret_cast<const uint4*>(a.weight) + i);
    if (tid < 32) {
        float sc = 1.f, sh = 0.f;
        if (a.x_mode == AARMVS_SRC_GN_RELU) {
            float mean, rstd;
            gn_mean_rstd(a.x_stats + (tid / (32 / a.x_groups)) * 2, a.x_count, mean, rstd);
            sc = rstd * a.x_gamma[tid];
            sh = a.x_beta[tid] - mean * sc;
        }
        s_scale[tid] = sc;
        s_shift[tid] = sh;
        s_bias[tid] = a.bias[tid];
    }
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

    const int tiles_x = (a.W + TILE_C - 1) / TILE_C;
    const int ntiles = tiles_x * ((a.H + TILE_R - 1) / TILE_R);
    const int my_tiles = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0) {
        // =========================================== MMA issuer ===========================================
        const bool leader = tc::elect_one();
        const uint32_t idesc2 = tc::idesc_f16_f32(128, 64), idesc1 = tc::idesc_f16_f32(128, 32);
        const uint32_t a_base = tc::smem_u32(sA), b_base = tc::smem_u32(sB);
        constexpr uint32_t A_HI = tc::desc_hi(8 * 16), B_HI = tc::desc_hi(8 * 16);
        const uint32_t b_lo = tc::desc_lo(b_base, 64 * 16);
        // group g owns this CTA's tile slots g, g + G, ...; its fills (tile, tap) arrive in order, the groups in lock step
        int fills[G];
        int rounds = 0;
        for (int g = 0; g < G; ++g) {
            fills[g] = ((my_tiles - g + G - 1) / G) * 9;
            rounds = max(rounds, fills[g]);
        }
        for (int r = 0; r < rounds; ++r) {
            const int tap = r % 9;
#pragma unroll
            for (int g = 0; g < G; ++g) {
                if (r >= fills[g]) continue;
                tc::mbar_wait(&full_a[g], r & 1);
                tc::tc_fence_after();
                const uint32_t a_lo_hi = tc::desc_lo(a_base + g * A_STAGE, AROWS * 16);
                const uint32_t a_lo_lo = tc::desc_lo(a_base + g * A_STAGE + A_PLANE, AROWS * 16);
                if (leader) {
                    const uint32_t dst = tmem_base + g * 64;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t ak = (uint32_t)(2 * ks * AROWS), bk = (uint32_t)((tap * 2 + ks) * (STEP_BYTES / 16));
                        tc::mma_f16_parts(dst, a_lo_hi + ak, A_HI, b_lo + bk, B_HI, idesc2, (tap | ks) != 0);
                        tc::mma_f16_parts(dst, a_lo_lo + ak, A_HI, b_lo + bk, B_HI, idesc1, 1);
                    }
                    tc::mma_commit(&done_a[g]);        // operand stage free; after tap 8: accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // =========================================== worker groups ===========================================
        const int g = (warp - 1) / 4, lt = tid - 32 - g * GROUP_THREADS;
        const int quad = warp & 3;                                   // TMEM lane quadrant this warp may read
        unsigned char* stage = sA + g * A_STAGE;
        Rec* recs = reinterpret_cast<Rec*>(sR + g * REC_BYTES);
        const uint32_t bar_id = 1 + g;
        const int c4i = lt & 7, px0 = lt >> 3;                        // this thread's 4-channel slice; item j is pixel px0 + 16 j
        const uint32_t st_thread = tc::smem_u32(stage) + (uint32_t)((c4i >> 1) * AROWS * 16 + (c4i & 1) * 8);
        const bool norm = a.x_mode == AARMVS_SRC_GN_RELU;
        const float4 sc = make_float4(s_scale[c4i * 4], s_scale[c4i * 4 + 1], s_scale[c4i * 4 + 2], s_scale[c4i * 4 + 3]);
        const float4 sh = make_float4(s_shift[c4i * 4], s_shift[c4i * 4 + 1], s_shift[c4i * 4 + 2], s_shift[c4i * 4 + 3]);
        const unsigned char* xb = reinterpret_cast<const unsigned char*>(a.x) + c4i * 16;
        const float hmax = (float)(a.H + 1), wmax = (float)(a.W + 1);      // padded size - 1 (module.py:153-159)
        float gs[4] = {0.f, 0.f, 0.f, 0.f}, gq[4] = {0.f, 0.f, 0.f, 0.f};
        int n_it = 0;
        for (int tl = g; tl < my_tiles; tl += G) {
          const int tile = blockIdx.x + tl * gridDim.x;
          const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
          const int pi = ty * TILE_R + (lt >> 3), pj = tx * TILE_C + (lt & 7);    // this thread's pixel for the recipes
          const bool valid = pi < a.H && pj < a.W;
          float om[27];
          {
              const float* q = a.offmask + (valid ? (size_t)(pi * a.W + pj) * 32 : 0);
#pragma unroll
              for (int c = 0; c < 24; c += 4) {
                  const float4 t = ldg4(q + c);
                  om[c] = t.x; om[c + 1] = t.y; om[c + 2] = t.z; om[c + 3] = t.w;
              }
              const float4 t = ldg4(q + 24);
              om[24] = t.x; om[25] = t.y; om[26] = t.z;
          }
#pragma unroll 1
          for (int k = 0; k < 9; ++k, ++n_it) {
            // ---- 1. recipe of pixel lt for tap k (module.py:150-165, 201-219: positions in the zero-padded image)
            {
                Rec rec;
                rec.o[0] = rec.o[1] = rec.o[2] = rec.o[3] = 0u;
                rec.g[0] = rec.g[1] = rec.g[2] = rec.g[3] = 0.f;
                rec.m = 0.f; rec.pad0 = rec.pad1 = rec.pad2 = 0.f;
                if (valid) {
                    float o_r = 0.f, o_c = 0.f, o_m = 0.f;
#pragma unroll
                    for (int t = 0; t < 9; ++t)
                        if (t == k) { o_r = om[t]; o_c = om[9 + t]; o_m = om[18 + t]; }      // registers, no dynamic indexing
                    const float pr_raw = __fadd_rn(__fadd_rn((float)(pi + 1), (float)(k / 3 - 1)), o_r);
                    const float pc_raw = __fadd_rn(__fadd_rn((float)(pj + 1), (float)(k % 3 - 1)), o_c);
                    rec.m = sigmoidf_acc(o_m);
                    const float fr = floorf(pr_raw), fc = floorf(pc_raw);
                    const float r0 = fminf(fmaxf(fr, 0.f), hmax), r1 = fminf(fmaxf(fr + 1.f, 0.f), hmax);
                    const float c0 = fminf(fmaxf(fc, 0.f), wmax), c1 = fminf(fmaxf(fc + 1.f, 0.f), wmax);
                    const float pr = fminf(fmaxf(pr_raw, 0.f), hmax), pc = fminf(fmaxf(pc_raw, 0.f), wmax);
                    // module.py:162-165 (weights from the CLAMPED corners and position)
                    const float g_lt = __fmul_rn(__fadd_rn(1.f, __fsub_rn(r0, pr)), __fadd_rn(1.f, __fsub_rn(c0, pc)));
                    const float g_rb = __fmul_rn(__fsub_rn(1.f, __fsub_rn(r1, pr)), __fsub_rn(1.f, __fsub_rn(c1, pc)));
                    const float g_lb = __fmul_rn(__fadd_rn(1.f, __fsub_rn(r0, pr)), __fsub_rn(1.f, __fsub_rn(c1, pc)));
                    const float g_rt = __fmul_rn(__fsub_rn(1.f, __fsub_rn(r1, pr)), __fadd_rn(1.f, __fsub_rn(c0, pc)));
                    // padded coordinates -> unpadded; the padding ring (index 0 and size + 1) holds zeros
                    const int ir0 = (int)r0 - 1, ir1 = (int)r1 - 1, ic0 = (int)c0 - 1, ic1 = (int)c1 - 1;
                    const bool vr0 = ir0 >= 0 && ir0 < a.H, vr1 = ir1 >= 0 && ir1 < a.H;
                    const bool vc0 = ic0 >= 0 && ic0 < a.W, vc1 = ic1 >= 0 && ic1 < a.W;
                    if (vr0 && vc0) { rec.o[0] = (unsigned)(ir0 * a.W + ic0) * 128u; rec.g[0] = g_lt; }
                    if (vr1 && vc1) { rec.o[1] = (unsigned)(ir1 * a.W + ic1) * 128u; rec.g[1] = g_rb; }
                    if (vr0 && vc1) { rec.o[2] = (unsigned)(ir0 * a.W + ic1) * 128u; rec.g[2] = g_lb; }
                    if (vr1 && vc0) { rec.o[3] = (unsigned)(ir1 * a.W + ic0) * 128u; rec.g[3] = g_rt; }
                }
                recs[lt] = rec;
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(GROUP_THREADS) : "memory");
            // ---- 2. gather, GroupNorm + ReLU of the input, blend (module.py:174-177, 184), split, stage
#pragma unroll 1
            for (int jb = 0; jb < 8; jb += 2) {
                float4 xa[2], xr[2], xl[2], xt[2], gw[2];
                float mm[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const Rec* rec = recs + px0 + (jb + j) * 16;
                    const uint4 o = *reinterpret_cast<const uint4*>(rec->o);
                    gw[j] = *reinterpret_cast<const float4*>(rec->g);
                    mm[j] = rec->m;
                    xa[j] = __ldg(reinterpret_cast<const float4*>(xb + o.x));
                    xr[j] = __ldg(reinterpret_cast<const float4*>(xb + o.y));
                    xl[j] = __ldg(reinterpret_cast<const float4*>(xb + o.z));
                    xt[j] = __ldg(reinterpret_cast<const float4*>(xb + o.w));
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    auto act = [&](float4 v) {
                        if (!norm) return v;
                        return make_float4(fmaxf(fmaf(v.x, sc.x, sh.x), 0.f), fmaxf(fmaf(v.y, sc.y, sh.y), 0.f),
                                           fmaxf(fmaf(v.z, sc.z, sh.z), 0.f), fmaxf(fmaf(v.w, sc.w, sh.w), 0.f));
                    };
                    const float4 A = act(xa[j]), R = act(xr[j]), L = act(xl[j]), T = act(xt[j]);
                    // a corner on the padding ring has weight 0 (and reads line 0): 0 * finite = 0, the reference's zero
                    auto blend = [&](float x_lt, float x_rb, float x_lb, float x_rt) {
                        const float s = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(gw[j].x, x_lt), __fmul_rn(gw[j].y, x_rb)),
                                                            __fmul_rn(gw[j].z, x_lb)), __fmul_rn(gw[j].w, x_rt));
                        return __fmul_rn(s, mm[j]);
                    };
                    const float4 v = make_float4(blend(A.x, R.x, L.x, T.x), blend(A.y, R.y, L.y, T.y), blend(A.z, R.z, L.z, T.z),
                                                 blend(A.w, R.w, L.w, T.w));
                    uint2 hi, lo;
                    tc::split4_f16(v, hi, lo);
                    const uint32_t addr = st_thread + (uint32_t)(px0 + (jb + j) * 16) * 16;
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(hi.x), "r"(hi.y) : "memory");
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr + A_PLANE), "r"(lo.x), "r"(lo.y) : "memory");
                }
            }
            tc::fence_async_smem();
            tc::mbar_arrive(&full_a[g]);
            // ---- 3. the stage (and, after tap 8, the tile's accumulator) is released by the MMAs of this fill
            tc::mbar_wait_relaxed(&done_a[g], n_it & 1);
          }
          // ---- 4. all nine taps accumulated: accumulator row = pixel quad * 32 + lane of the tile
          {
            tc::tc_fence_after();
            float acc[32], acc2[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + g * 64;
            tc::tmem_ld32(taddr, acc);
            tc::tmem_ld32(taddr + 32, acc2);
            tc::tc_fence_before();                                       // ordered before this thread's next arrive on full[g]
            const int m = quad * 32 + lane, oy = ty * TILE_R + (m >> 3), ox = tx * TILE_C + (m & 7);
            if (oy < a.H && ox < a.W) {
                float* o = a.out + (size_t)(oy * a.W + ox) * 32;
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float4 r = make_float4((acc[c] + acc2[c]) + s_bias[c], (acc[c + 1] + acc2[c + 1]) + s_bias[c + 1],
                                                 (acc[c + 2] + acc2[c + 2]) + s_bias[c + 2], (acc[c + 3] + acc2[c + 3]) + s_bias[c + 3]);
                    st4(o + c, r);
                    gs[c / 8] += (r.x + r.y) + (r.z + r.w);
                    gq[c / 8] += (r.x * r.x + r.y * r.y) + (r.z * r.z + r.w * r.w);
                }
            }
          }
        }
        if (a.out_stats) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double s = warp_sum((double)gs[q]), qq = warp_sum((double)gq[q]);
                if (lane == 0) {
                    atomicAdd(a.out_stats + q * 2 + 0, s);
                    atomicAdd(a.out_stats + q * 2 + 1, qq);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace

extern "C" int aarmvs_deform_conv_tc(const aarmvs_conv_src* x, int H, int W, const float* offmask, const void* weight_f16,
                                     const float* bias, float* out, double* out_gn_sums, void* stream) {
    AARMVS_CHECK_ARG(x && x->ptr && offmask && weight_f16 && bias && out && H > 0 && W > 0 && x->channels == 32);
    AARMVS_CHECK_ARG(x->mode == AARMVS_SRC_PLAIN || x->mode == AARMVS_SRC_GN_RELU);
    AARMVS_CHECK_ARG(((uintptr_t)weight_f16 & 15) == 0 && (size_t)H * W * 128 < 0xFFFFFFFFull);
    if (x->mode == AARMVS_SRC_GN_RELU)
        AARMVS_CHECK_ARG(x->gn_sums && x->gn_weight && x->gn_bias && x->gn_groups >= 1 && 32 % x->gn_groups == 0 && x->gn_count > 0);
    DeformTcArgs a = {};
    a.x = x->ptr; a.x_stats = x->gn_sums; a.x_gamma = x->gn_weight; a.x_beta = x->gn_bias; a.x_count = x->gn_count;
    a.x_groups = x->gn_groups; a.x_mode = x->mode; a.H = H; a.W = W; a.offmask = offmask; a.weight = (const __half*)weight_f16;
    a.bias = bias; a.out = out; a.out_stats = out_gn_sums;
    static unsigned long long smem_ok = 0;
    if (cudaError_t e = ensure_dyn_smem(deform_tc_kernel, SMEM, smem_ok)) return (int)e;
    const int ntiles = ((W + TILE_C - 1) / TILE_C) * ((H + TILE_R - 1) / TILE_R);
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    deform_tc_kernel<<<ntiles < sms ? ntiles : sms, N_THREADS, SMEM, (cudaStream_t)stream>>>(a);
    AARMVS_LAUNCH_RESULT();
}
