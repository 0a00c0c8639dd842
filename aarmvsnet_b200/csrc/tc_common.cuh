// tcgen05 / TMEM / mbarrier PTX wrappers (sm_100a) and the UMMA descriptor encodings used by the ConvLSTM
// implicit-GEMM kernels.  Encodings follow the PTX ISA tcgen05 matrix/instruction descriptor tables:
//   shared-memory matrix descriptor (64 bit):  [0,14) start address >> 4 | [16,30) leading byte offset >> 4 |
//       [32,46) stride byte offset >> 4 | [46,48) version = 1 | [49,52) base offset | [61,64) swizzle (0 = none)
//   K-major operand WITHOUT swizzle ("interleaved" canonical layout, units of 16 B):
//       ((8, m), 2) : ((1, SBO), LBO)     8 rows x 16 B contiguous = one 128-byte core matrix; the next 8-row group is
//       SBO bytes further; the second 16-byte K chunk of the K=16 (fp16) instruction is LBO bytes further.
//   instruction descriptor (32 bit, kind::f16): [4,6) D format (1 = F32) | [7,10) A format (0 = F16) | [10,13) B format |
//       bit 15 / 16 A / B major (0 = K) | [17,23) N >> 3 | [24,29) M >> 4
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc {

// Programmatic dependent launch (PTX griddepcontrol).  The kernels of the per-plane recurrence chain are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (common.cuh: launch_pdl): a kernel's CTAs may become resident while its
// predecessor drains, run the part of their prologue that touches nothing a predecessor writes (barrier init, TMEM
// allocation, tensor-map prefetch, the bulk load of the immutable weights), and only then wait for the predecessor's
// completion (and memory flush).  launch_dependents is issued right AFTER the wait, so at most two kernels of the chain
// overlap and "complete" is transitive.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__host__ __device__ constexpr uint32_t idesc_f16_f32(int M, int N) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- mbarrier -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch, never as a hung GPU.
// Latency-critical waiter (the MMA-issuing thread): tight polling.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 28); ++spin)
        if (mbar_try_wait(bar, parity)) return;
    __trap();
}
// Everyone else backs off between polls so that waiting warps do not take issue slots from working ones
// (ncu: 55 % of the executed instructions of the first version were barrier polling).
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
#ifdef AARMVS_WAIT_NANOSLEEP
    if (mbar_try_wait(bar, parity)) return;
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        __nanosleep(AARMVS_WAIT_NANOSLEEP);
        if (mbar_try_wait(bar, parity)) return;
    }
    __trap();
#else
    mbar_wait(bar, parity);          // try_wait already suspends the warp in hardware for a bounded time per poll
#endif
}

// ---- TMA (cp.async.bulk[.tensor]) --------------------------------------------------------------------------
// One arrival + `bytes` expected from the bulk copies that name this barrier.
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 5-D tiled tensor load global -> shared (box and strides live in the CUtensorMap); out-of-bounds elements are
// written as zeros, which is exactly the conv's zero padding.  dst must be 128-byte aligned.
__device__ __forceinline__ void tma_load_5d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// Linear bulk copy global -> shared; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {       // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {         // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], one elected thread
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same, with the descriptors given as 32-bit halves: the high words (SBO, version) are per-kernel constants and the low
// words are "base + immediate", so an unrolled tap loop costs one 32-bit add per operand instead of rebuilding 64-bit
// descriptors (the single MMA-issuing thread was the serial bottleneck of the first version: ~12 instructions per MMA).
__device__ __forceinline__ uint32_t desc_lo(uint32_t addr, uint32_t lbo_bytes) { return ((addr >> 4) & 0x3FFF) | (((lbo_bytes >> 4) & 0x3FFF) << 16); }
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14); }
__device__ __forceinline__ void mma_f16_parts(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// all previously issued MMAs of this thread complete -> one arrival on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives row (lane base + i), columns [col, col+32)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Gate non-linearities for the tensor-core epilogues: MUFU.EX2 + MUFU.RCP (abs error ~1e-7, the same order as the
// fp32 rounding of the reference's own sigmoid / tanh), 4-5 instructions instead of ~12 / ~25 for expf-div / tanhf.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_sigmoid(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }
__device__ __forceinline__ float fast_tanh(float x) {       // 2 / (1 + e^{-2x}) - 1
    return fmaf(2.0f, rcp_approx(1.0f + ex2_approx(-2.8853900817779268f * x)), -1.0f);
}

// fp32 -> (hi, lo) fp16 pair with x ~= hi + lo to ~2^-22 relative (|x| < 65504; lo may be subnormal: abs error <= 2^-25)
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
    hi = __float2half_rn(x);
    lo = __float2half_rn(x - __half2float(hi));
}

// One value, the same truncating split as split4_f16.
__device__ __forceinline__ void split_f16_trunc(float x, __half& hi, __half& lo) {
    const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    hi = __float2half_rn(h);
    lo = __float2half_rn(x - h);
}

// Four fp32 values -> packed fp16 (hi) and (lo) pairs.  hi = x with the low 13 mantissa bits cleared (exactly an fp16
// value in the normal range), lo = x - hi (exact in fp32, <= 2^-11 |x|, rounded to fp16): x ~= hi + lo to 2^-22 |x|.
// 3.5 instructions per element (LOP, FADD, 2 x half of a packed F2F) instead of 5+ for round-to-nearest splitting.
__device__ __forceinline__ void split4_f16(float4 v, uint2& hi, uint2& lo) {
    const float h0 = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u), h1 = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
    const float h2 = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u), h3 = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
    __half2 a = __floats2half2_rn(h0, h1), b = __floats2half2_rn(h2, h3);
    __half2 c = __floats2half2_rn(v.x - h0, v.y - h1), d = __floats2half2_rn(v.z - h2, v.w - h3);
    hi = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    lo = make_uint2(*reinterpret_cast<uint32_t*>(&c), *reinterpret_cast<uint32_t*>(&d));
}

}  // namespace tc
