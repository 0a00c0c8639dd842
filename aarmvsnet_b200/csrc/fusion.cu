// Depth-map fusion filter of one reference view (fusion.py:71-133 reproject_with_depth /
// check_geometric_consistency and the accumulation of filter_depth, fusion.py:179-225): one thread per reference pixel
// walks the source views -- project with the reference depth, sample the source depth map (cv2.remap's 1/32-pixel
// fixed-point bilinear, border 0), project back, test the nine nested (distance, relative depth) thresholds -- and keeps
// every per-view intermediate in registers (the reference materialises ~20 full-size float64 arrays per source view).
// Arithmetic follows numpy's dtype flow: float64 geometry from float32 camera matrices (prepared on the host),
// float32 roundings exactly where the reference casts (.astype(np.float32)), float32 relative-depth test.
#include "common.cuh"
#include "../../include/aarmvs.h"

namespace {

constexpr int MATS_PER_VIEW = 60;      // Kr_inv[9] T_rs[12] Ks[9] Ks_inv[9] T_sr[12] Kr[9]  (row-major; T: top 3 rows of the 4x4)

__device__ __forceinline__ void mat3(const double* __restrict__ m, double x, double y, double z, double& ox, double& oy, double& oz) {
    ox = m[0] * x + m[1] * y + m[2] * z;
    oy = m[3] * x + m[4] * y + m[5] * z;
    oz = m[6] * x + m[7] * y + m[8] * z;
}
__device__ __forceinline__ void mat34(const double* __restrict__ m, double x, double y, double z, double& ox, double& oy, double& oz) {
    ox = m[0] * x + m[1] * y + m[2] * z + m[3];
    oy = m[4] * x + m[5] * y + m[6] * z + m[7];
    oz = m[8] * x + m[9] * y + m[10] * z + m[11];
}

// cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) for a float32 image: OpenCV rounds the map to 1/32 pixel (cvRound: to nearest
// even), takes the weights from its float table (1 - a/32, a/32) and sums the four products left to right without FMA.
__device__ __forceinline__ float remap_bilinear(const float* __restrict__ img, int H, int W, float mx, float my) {
    const float fx = rintf(__fmul_rn(mx, 32.0f)), fy = rintf(__fmul_rn(my, 32.0f));
    if (!(fabsf(fx) <= 2.0e9f) || !(fabsf(fy) <= 2.0e9f)) return 0.0f;          // NaN / overflow: cvRound gives INT_MIN -> outside
    const int sx = (int)fx, sy = (int)fy;
    const int ix = max(-32768, min(32767, sx >> 5)), iy = max(-32768, min(32767, sy >> 5));
    const float ax = __fmul_rn((float)(sx & 31), 0.03125f), ay = __fmul_rn((float)(sy & 31), 0.03125f);
    const float wx0 = __fsub_rn(1.0f, ax), wy0 = __fsub_rn(1.0f, ay);
    auto tap = [&](int yy, int xx) { return (xx >= 0 && xx < W && yy >= 0 && yy < H) ? __ldg(img + (size_t)yy * W + xx) : 0.0f; };
    const float v00 = tap(iy, ix), v01 = tap(iy, ix + 1), v10 = tap(iy + 1, ix), v11 = tap(iy + 1, ix + 1);
    float r = __fmul_rn(v00, __fmul_rn(wy0, wx0));
    r = __fadd_rn(r, __fmul_rn(v01, __fmul_rn(wy0, ax)));
    r = __fadd_rn(r, __fmul_rn(v10, __fmul_rn(ay, wx0)));
    r = __fadd_rn(r, __fmul_rn(v11, __fmul_rn(ay, ax)));
    return r;
}

__global__ void __launch_bounds__(256) fuse_view_kernel(const float* __restrict__ depth_ref, const float* __restrict__ conf_ref,
                                                        const float* __restrict__ depth_src, int S, int H, int W,
                                                        const double* __restrict__ mats, float photo_threshold,
                                                        unsigned char* __restrict__ geo_mask, unsigned char* __restrict__ final_mask,
                                                        float* __restrict__ depth_avg, int* __restrict__ geo_sum_out) {
    const size_t P = (size_t)H * W;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
        const int y = (int)(p / W), x = (int)(p - (size_t)y * W);
        const float dref = depth_ref[p];
        int count[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};       // views passing threshold i = 2..10 (fusion.py:127-129)
        float acc = 0.0f;                                  // sum(all_srcview_depth_ests): float32, in view order
        for (int v = 0; v < S; ++v) {
            const double* m = mats + (size_t)v * MATS_PER_VIEW;
            const double dd = (double)dref;
            double X, Y, Z, xs, ys, zs;
            mat3(m, (double)x * dd, (double)y * dd, dd, X, Y, Z);                    // reference camera space
            mat34(m + 9, X, Y, Z, xs, ys, zs);                                       // source camera space
            double kx, ky, kz;
            mat3(m + 21, xs, ys, zs, kx, ky, kz);
            const double u = kx / kz, w = ky / kz;                                   // source pixel
            const float sampled = remap_bilinear(depth_src + (size_t)v * P, H, W, (float)u, (float)w);
            const double sd = (double)sampled;
            mat3(m + 30, u * sd, w * sd, sd, X, Y, Z);                               // source space at the SOURCE view's depth
            double xr, yr, zr;
            mat34(m + 39, X, Y, Z, xr, yr, zr);                                      // back in the reference camera space
            const float depth_rep = (float)zr;
            mat3(m + 51, xr, yr, zr, kx, ky, kz);
            const float x_rep = (float)(kx / kz), y_rep = (float)(ky / kz);
            const double ex = (double)x_rep - (double)x, ey = (double)y_rep - (double)y;
            const double dist = sqrt(ex * ex + ey * ey);
            const float rel = __fdiv_rn(fabsf(__fsub_rn(depth_rep, dref)), dref);
            bool last = false;
#pragma unroll
            for (int i = 2; i <= 10; ++i) {
                const bool ok = dist < (double)i / 4.0 && rel < (float)((double)i / 1300.0);
                count[i - 2] += ok ? 1 : 0;
                last = ok;
            }
            acc = __fadd_rn(acc, last ? depth_rep : 0.0f);                           // depth_reprojected[~mask] = 0
        }
        const int n = S + 1, gsum = count[8];
        bool geo = gsum >= n;                                                        // never true: kept for fidelity (fusion.py:219)
#pragma unroll
        for (int i = 2; i <= 10; ++i)
            if (i < n) geo = geo || count[i - 2] >= i;
        // (sum + ref_depth) is float32, the division by the int32 count promotes to float64 in numpy
        const float avg = (float)((double)__fadd_rn(acc, dref) / (double)(gsum + 1));
        geo_mask[p] = geo ? 1 : 0;
        if (final_mask) final_mask[p] = (geo && conf_ref && conf_ref[p] > photo_threshold) ? 1 : 0;
        depth_avg[p] = avg;
        if (geo_sum_out) geo_sum_out[p] = gsum;
    }
}

}  // namespace

extern "C" int aarmvs_fuse_view_mats(void) { return MATS_PER_VIEW; }

extern "C" int aarmvs_fuse_view(const float* depth_ref, const float* conf_ref, const float* depth_src, int S, int H, int W,
                                const double* mats, float photo_threshold, unsigned char* geo_mask, unsigned char* final_mask,
                                float* depth_avg, int* geo_sum, void* stream) {
    AARMVS_CHECK_ARG(depth_ref && depth_src && mats && geo_mask && depth_avg && H > 0 && W > 0);
    AARMVS_CHECK_ARG(S >= 1 && S <= 10);              // the reference keeps nine nested thresholds: at most 10 source views
    AARMVS_CHECK_ARG(!final_mask || conf_ref);
    const size_t P = (size_t)H * W;
    const int sms = aarmvs_sm_count();
    if (sms <= 0) return AARMVS_ERR_ARG;
    size_t blocks = (P + 255) / 256;
    if (blocks > (size_t)sms * 8) blocks = (size_t)sms * 8;
    fuse_view_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(depth_ref, conf_ref, depth_src, S, H, W, mats,
                                                                         photo_threshold, geo_mask, final_mask, depth_avg, geo_sum);
    AARMVS_LAUNCH_RESULT();
}
