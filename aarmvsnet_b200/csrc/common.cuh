// Shared device helpers for the aarmvs sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#define AARMVS_OK 0
#define AARMVS_ERR_ARG 1000001   // bad argument (shape / alignment / null pointer)

#define AARMVS_CHECK_ARG(cond) do { if (!(cond)) return AARMVS_ERR_ARG; } while (0)
#define AARMVS_LAUNCH_RESULT() do { cudaError_t e__ = cudaGetLastError(); return e__ == cudaSuccess ? AARMVS_OK : (int)e__; } while (0)

constexpr int AARMVS_MAX_SRC = 15;     // source views per sample (N - 1)
constexpr int AARMVS_MAX_PLANES = 16;  // depth planes per launch of the plane-batched passes A-D
constexpr float GN_EPS = 1e-5f;        // nn.GroupNorm default (models/module.py:101)

struct Cams {                          // per source view: rot (row-major 3x3) then trans (3), models/module.py:16-18
    float m[AARMVS_MAX_SRC][12];
};

// Opt-in to > 48 KB of dynamic shared memory once per (kernel, device); `mask` is the call site's static bit set.
template <class K>
static inline cudaError_t ensure_dyn_smem(K kern, int bytes, unsigned long long& mask) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && ((mask >> dev) & 1ull)) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev < 64) mask |= 1ull << dev;
    return e;
}

// Grid of a persistent kernel with balanced rounds: the fewest CTAs that still finish in ceil(tiles / sms) rounds.  160 tiles on
// 148 SMs: 80 CTAs x 2 tiles instead of 148 CTAs of which 12 take two -- the same latency on half the SMs, which the concurrent
// streams of a training step use; large maps (thousands of tiles) keep one CTA per SM.
static inline int balanced_grid(int tiles, int sms) {
    if (tiles <= sms) return tiles;
    const int rounds = (tiles + sms - 1) / sms;
    return (tiles + rounds - 1) / rounds;
}

// 16-byte asynchronous global -> shared copy (no register round trip: a staging loop keeps all its loads in flight);
// !pred writes 16 zero bytes instead (src is not read, but must be a valid address).
__device__ __forceinline__ void cp_async16_zfill(float* smem_dst, const float* src, bool pred) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Launch with programmatic stream serialization (see tc_common.cuh: pdl_wait).  AARMVS_NO_PDL=1 in the environment falls back
// to a plain launch (A/B measurements).
template <class Arg>
static inline cudaError_t launch_pdl(void (*kern)(Arg), dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Arg& arg) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    const char* off = getenv("AARMVS_NO_PDL");
    cfg.attrs = attr;
    cfg.numAttrs = (off && off[0] == '1') ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, arg);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ int plane_index(int d_host, const int* d_dev) { return d_dev ? *d_dev : d_host; }

// mean / rstd of a GroupNorm group from its fp64 running sums (biased variance, eps inside the sqrt).
__device__ __forceinline__ void gn_mean_rstd(const double* s, double count, float& mean, float& rstd) {
    double m = s[0] / count;
    double var = s[1] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)GN_EPS));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of NV doubles per thread; thread 0 of the block receives the totals in out[].
template <int NV, int NWARPS>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* NV * NWARPS doubles */, double (&out)[NV]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = warp_sum(v[i]);
        if (lane == 0) smem[i * NWARPS + warp] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
            for (int w = 0; w < NWARPS; ++w) s += smem[i * NWARPS + w];
            out[i] = s;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------------------------------------------
// Plane-sweep homography (models/module.py:20-34) + grid_sample unnormalisation (align_corners switch).
// Arithmetic order mirrors the reference op by op so the sample position is bit-identical in fp32:
//   rot @ (x, y, 1): fma chain over k = 0, 1, 2 (what ATen's K=3 sgemm does);  * depth;  + trans;
//   z == 0 -> z += 1e-4;  x / z, y / z;  / ((W - 1) / 2) - 1;  then ((g + 1) * W - 1) / 2  (align_corners=False).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sweep_position(const float* cam, float x, float y, float depth, int W, int H,
                                               bool align_corners, float& ix, float& iy) {
    float a0 = __fmul_rn(cam[0], x); a0 = __fmaf_rn(cam[1], y, a0); a0 = __fadd_rn(a0, cam[2]);
    float a1 = __fmul_rn(cam[3], x); a1 = __fmaf_rn(cam[4], y, a1); a1 = __fadd_rn(a1, cam[5]);
    float a2 = __fmul_rn(cam[6], x); a2 = __fmaf_rn(cam[7], y, a2); a2 = __fadd_rn(a2, cam[8]);
    float px = __fadd_rn(__fmul_rn(a0, depth), cam[9]);
    float py = __fadd_rn(__fmul_rn(a1, depth), cam[10]);
    float pz = __fadd_rn(__fmul_rn(a2, depth), cam[11]);
    if (pz == 0.0f) pz = __fadd_rn(pz, 0.0001f);
    float qx = __fdiv_rn(px, pz), qy = __fdiv_rn(py, pz);
    float gx = __fadd_rn(__fdiv_rn(qx, (float)(W - 1) * 0.5f), -1.0f);
    float gy = __fadd_rn(__fdiv_rn(qy, (float)(H - 1) * 0.5f), -1.0f);
    // ATen's CPU grid sampler (GridSamplerKernel.cpp, ComputeLocation::unnormalize) evaluates
    // (g + 1) * scale - 0.5 with the multiply-add contracted to one FMA; verified bit-exact against
    // F.grid_sample on CPU ("bit-exact sample positions").
    if (align_corners) {
        ix = __fmul_rn(__fadd_rn(gx, 1.0f), (float)(W - 1) * 0.5f);
        iy = __fmul_rn(__fadd_rn(gy, 1.0f), (float)(H - 1) * 0.5f);
    } else {
        ix = __fmaf_rn(__fadd_rn(gx, 1.0f), (float)W * 0.5f, -0.5f);
        iy = __fmaf_rn(__fadd_rn(gy, 1.0f), (float)H * 0.5f, -0.5f);
    }
}

// Bilinear tap set of one sample position: 4 corner offsets (in pixels, -1 = out of bounds -> zero) + weights.
struct Taps {
    int o_nw, o_ne, o_sw, o_se;
    float w_nw, w_ne, w_sw, w_se;
    int xi, yi;                      // integer coordinates of the north-west corner (0 when no corner can be inside)
};

__device__ __forceinline__ Taps make_taps(float ix, float iy, int W, int H) {
    Taps t;
    float fx = floorf(ix), fy = floorf(iy);
    float ex = __fadd_rn(fx, 1.0f), sy = __fadd_rn(fy, 1.0f);
    // distances to the west / east / north / south neighbours, then weights (ATen compute_interp_params)
    float dw = __fsub_rn(ix, fx), de = __fsub_rn(1.0f, dw), dn = __fsub_rn(iy, fy), ds = __fsub_rn(1.0f, dn);
    t.w_nw = __fmul_rn(ds, de);
    t.w_ne = __fmul_rn(ds, dw);
    t.w_sw = __fmul_rn(dn, de);
    t.w_se = __fmul_rn(dn, dw);
    // bounds in the float domain (NaN / huge coordinates fail every test and contribute zero)
    bool x0 = fx >= 0.0f && fx <= (float)(W - 1), x1 = ex >= 0.0f && ex <= (float)(W - 1);
    bool y0 = fy >= 0.0f && fy <= (float)(H - 1), y1 = sy >= 0.0f && sy <= (float)(H - 1);
    int xi = x0 || x1 ? (int)fx : 0, yi = y0 || y1 ? (int)fy : 0;
    t.xi = xi; t.yi = yi;
    t.o_nw = (x0 && y0) ? yi * W + xi : -1;
    t.o_ne = (x1 && y0) ? yi * W + xi + 1 : -1;
    t.o_sw = (x0 && y1) ? (yi + 1) * W + xi : -1;
    t.o_se = (x1 && y1) ? (yi + 1) * W + xi + 1 : -1;
    return t;
}

// 4 channels (this lane's chunk) of the bilinear sample; src is NHWC with 32 channels.
__device__ __forceinline__ float4 gather_taps(const float* __restrict__ src, const Taps& t, int chunk) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 a = t.o_nw >= 0 ? ldg4(src + (size_t)t.o_nw * 32 + chunk * 4) : z;
    float4 b = t.o_ne >= 0 ? ldg4(src + (size_t)t.o_ne * 32 + chunk * 4) : z;
    float4 c = t.o_sw >= 0 ? ldg4(src + (size_t)t.o_sw * 32 + chunk * 4) : z;
    float4 d = t.o_se >= 0 ? ldg4(src + (size_t)t.o_se * 32 + chunk * 4) : z;
    float4 r;
    // fma(se, w, fma(sw, w, fma(ne, w, nw * w))): the contraction ATen's vectorised CPU kernel compiles to
    r.x = __fmaf_rn(d.x, t.w_se, __fmaf_rn(c.x, t.w_sw, __fmaf_rn(b.x, t.w_ne, __fmul_rn(a.x, t.w_nw))));
    r.y = __fmaf_rn(d.y, t.w_se, __fmaf_rn(c.y, t.w_sw, __fmaf_rn(b.y, t.w_ne, __fmul_rn(a.y, t.w_nw))));
    r.z = __fmaf_rn(d.z, t.w_se, __fmaf_rn(c.z, t.w_sw, __fmaf_rn(b.z, t.w_ne, __fmul_rn(a.z, t.w_nw))));
    r.w = __fmaf_rn(d.w, t.w_se, __fmaf_rn(c.w, t.w_sw, __fmaf_rn(b.w, t.w_ne, __fmul_rn(a.w, t.w_nw))));
    return r;
}

// Packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2): IEEE round-to-nearest per lane, i.e. bit-identical to the scalar
// __fmaf_rn / __fmul_rn / __fsub_rn chains they replace, at half the issue slots.
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// Gather recipe of pixel (x, y) in one source view at one depth: the four bilinear weights in tap order north-west, north-east,
// south-west, south-east and the byte offset `base` of the north-west line.  The taps are read at base, base + 128,
// base + 128 W and base + 128 W + 128 -- fixed steps, so the three neighbours are immediate / one-add addresses -- and a
// corner outside the image has weight 0, which reproduces ATen's skipped tap exactly (its line is still read: the caller
// guarantees (W + 1) * 128 readable, finite bytes behind the last source map).  On the low side nothing is read before a map:
// with the west (north) column (row) outside, the two valid taps move into the west (north) weight slots and the base is
// clamped to column (row) 0; the blend fma(se, w, fma(sw, w, fma(ne, w, nw * w))) only gains exact +-0 terms that way, the
// result is bit-identical.
__device__ __forceinline__ void make_recipe(const float* cam, int x, int y, float depth, int W, int H, bool align_corners,
                                            float4& wts, unsigned& base) {
    float ix, iy;
    sweep_position(cam, (float)x, (float)y, depth, W, H, align_corners, ix, iy);
    const Taps t = make_taps(ix, iy, W, H);
    float w0 = t.o_nw >= 0 ? t.w_nw : 0.f, w1 = t.o_ne >= 0 ? t.w_ne : 0.f, w2 = t.o_sw >= 0 ? t.w_sw : 0.f, w3 = t.o_se >= 0 ? t.w_se : 0.f;
    int xi = t.xi, yi = t.yi;                       // north-west corner: -1 .. size-1 on each axis (0 when no corner is inside)
    if (xi < 0) { w0 = w1; w1 = 0.f; w2 = w3; w3 = 0.f; xi = 0; }
    if (yi < 0) { w0 = w2; w1 = w3; w2 = 0.f; w3 = 0.f; yi = 0; }
    wts = make_float4(w0, w1, w2, w3);
    base = (unsigned)(yi * W + xi) * 128u;
}
