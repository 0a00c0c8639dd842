// Known-answer test of the tcgen05 building blocks used by the ConvLSTM kernels: un-swizzled K-major operands with
// a non-trivial stride between 8-row groups and a shifted start row (the "shifted window" addressing of a 3x3 tap
// inside a halo tile), TMEM accumulation over several K=16 instructions, commit -> mbarrier, tcgen05.ld read-back.
//   D[m][n] = sum_k A[shift + (m / 8) * group_rows + (m % 8)][k] * B[n][k],   m < 128, n < 64
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/aarmvs.h"

__global__ void __launch_bounds__(128) tc_selftest_kernel(const __half* __restrict__ A, int a_rows, const __half* __restrict__ B,
                                                          float* __restrict__ D, int K, int group_rows, int shift) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int kc = K / 8;                                  // 16-byte K chunks
    __half* sA = reinterpret_cast<__half*>(smem);          // [kc][a_rows][8]
    __half* sB = sA + (size_t)kc * a_rows * 8;             // [kc][64][8]
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kc * a_rows; i += 128) {
        const int c = i / a_rows, r = i - c * a_rows;
        *reinterpret_cast<uint4*>(sA + (size_t)i * 8) = *reinterpret_cast<const uint4*>(A + (size_t)r * K + c * 8);
    }
    for (int i = tid; i < kc * 64; i += 128) {
        const int c = i / 64, n = i - c * 64;
        *reinterpret_cast<uint4*>(sB + (size_t)i * 8) = *reinterpret_cast<const uint4*>(B + (size_t)n * K + c * 8);
    }
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc<64>(&tmem_base);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_f16_f32(128, 64);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t da = tc::smem_desc(tc::smem_u32(sA) + (uint32_t)((2 * ks) * a_rows + shift) * 16, (uint32_t)a_rows * 16,
                                              (uint32_t)group_rows * 16);
            const uint64_t db = tc::smem_desc(tc::smem_u32(sB) + (uint32_t)(2 * ks) * 64 * 16, 64 * 16, 8 * 16);
            tc::mma_f16(tmem, da, db, idesc, ks > 0);
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    float v[32];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + half * 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) D[(size_t)tid * 64 + half * 32 + j] = v[j];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<64>(tmem);
}

extern "C" int aarmvs_tc_selftest(const void* a_f16, int a_rows, const void* b_f16, float* d, int K, int group_rows,
                                  int shift, void* stream) {
    AARMVS_CHECK_ARG(a_f16 && b_f16 && d && K >= 16 && K % 16 == 0 && K <= 256 && group_rows >= 8 && shift >= 0);
    AARMVS_CHECK_ARG(shift + 15 * group_rows + 8 <= a_rows && a_rows <= 512);
    const size_t smem = (size_t)(K / 8) * (a_rows + 64) * 16;
    AARMVS_CHECK_ARG(smem <= 200 * 1024);
    cudaError_t e = cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __half*)a_f16, a_rows, (const __half*)b_f16, d, K, group_rows, shift);
    AARMVS_LAUNCH_RESULT();
}

// ---- micro-benchmark: cycles per tcgen05.mma (M = 128, K = 16, fp16) as a function of N and of the operand layout ----
__global__ void __launch_bounds__(128) tc_mma_bench_kernel(int N, int group_rows, int reps, int a_shift_rows, long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    if (warp == 0) tc::tmem_alloc<256>(&tmem_base);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    if (warp == 0) {
        const bool leader = tc::elect_one();
        const uint32_t idesc = tc::idesc_f16_f32(128, N);
        const uint32_t a_lo = tc::desc_lo(tc::smem_u32(smem), 196 * 16), b_lo = tc::desc_lo(tc::smem_u32(smem) + 24 * 1024, N * 16);
        const uint32_t a_hi = tc::desc_hi(group_rows * 16), b_hi = tc::desc_hi(8 * 16);
        long long t0 = clock64();
        if (leader) {
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int j = 0; j < 9; ++j) tc::mma_f16_parts(tmem_base, a_lo + j * a_shift_rows, a_hi, b_lo, b_hi, idesc, 1);
            }
            tc::mma_commit(&bar);
        }
        __syncwarp();
        tc::mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (leader) out[blockIdx.x] = t1 - t0;
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem_base);
}

// Debug: cycles for reps*9 MMAs (M=128, N, K=16) on every SM; out: [grid] int64.
extern "C" int aarmvs_tc_mma_bench(int N, int group_rows, int reps, int a_shift_rows, int grid, long long* out, void* stream) {
    AARMVS_CHECK_ARG(out && N >= 16 && N <= 256 && N % 16 == 0 && reps > 0 && grid > 0);
    cudaError_t e = cudaFuncSetAttribute(tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    if (e != cudaSuccess) return (int)e;
    tc_mma_bench_kernel<<<grid, 128, 48 * 1024, (cudaStream_t)stream>>>(N, group_rows, reps, a_shift_rows, out);
    AARMVS_LAUNCH_RESULT();
}
