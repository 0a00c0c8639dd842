/* aarmvs.h -- C ABI of libaarmvs_sm100.so: the sm_100a kernels of the AA-RMVSNet per-depth-plane cost-volume
 * pipeline.  The reference (BuTTerK3ks/AA-RMVSNet) is pure PyTorch and has no FFI of its own; each entry point
 * below replaces the reference Python call site cited next to it (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - activations are fp32 NHWC ("pixel major": [H][W][C]); images arrive as planar NCHW;
 *   - no allocation, no ownership transfer, no global mutable state: the caller owns every buffer and passes
 *     the CUDA stream (cudaStream_t as void*) the work is enqueued on; calls are asynchronous;
 *   - return value: 0 = enqueued; AARMVS_ERR_ARG = rejected argument; otherwise a cudaError_t value;
 *   - plane index: kernels of the per-plane chain take (d_host, d_dev).  If d_dev is non-null the plane index
 *     is read from device memory at run time, which lets one captured CUDA graph be replayed for every plane.
 *     Hidden states h live in ping-pong pairs: plane d writes half (d & 1) and reads h_{d-1} from the other.
 */
#ifndef AARMVS_H
#define AARMVS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define AARMVS_ERR_ARG 1000001
#define AARMVS_MAX_SRC_VIEWS 15
#define AARMVS_MAX_PLANES_PER_CALL 16

int aarmvs_version(void);          /* ABI version, currently 5 (5: aarmvs_convlstm_cell_tc gained gates_out; aarmvs_lstm_gates_bwd_act, aarmvs_conv_dx_tc, aarmvs_conv_dw_tc, aarmvs_split_planes_scaled / _cat) */
int aarmvs_sm_count(void);         /* multiprocessors of the current device (148 on B200), <0 on error */

/* ---------------------------------------------------------------------------------------------------------
 * Input description for the convolution kernels: a conv reads the channel concatenation of up to three
 * sources (replaces torch.cat at models/module.py:80 and models/drmvsnet.py:157,161) and can apply the
 * producer's GroupNorm+ReLU (models/module.py:101-102) or a 2x2/2 max-pool (models/drmvsnet.py:148,152) on load.
 */
enum { AARMVS_SRC_PLAIN = 0, AARMVS_SRC_GN_RELU = 1, AARMVS_SRC_MAXPOOL2 = 2, AARMVS_SRC_NCHW = 3 };

typedef struct aarmvs_conv_src {
    const float* ptr;          /* NHWC [H][W][channels]; MAXPOOL2: [2H][2W][channels]; NCHW: planar [planes][H][W] */
    int channels;              /* channels contributed to the concat (multiple of 4; NCHW: planes padded up to 4) */
    int mode;                  /* AARMVS_SRC_* */
    int planes;                /* NCHW only: real plane count (3 for RGB) */
    int gn_groups;             /* GN_RELU only */
    const double* gn_sums;     /* GN_RELU only: [gn_groups][2] = (sum, sum of squares) written by the producer */
    const float* gn_weight;    /* GN_RELU only: [channels] */
    const float* gn_bias;      /* GN_RELU only: [channels] */
    double gn_count;           /* GN_RELU only: elements per group */
    size_t pp_stride;          /* 0, or floats between the two halves of a ping-pong pair */
    int pp_phase;              /* with pp_stride: 0 = half written by the current plane (d & 1), 1 = previous plane */
} aarmvs_conv_src;

/* ---- FeatNet (models/drmvsnet.py:41-63, 7-24) ---------------------------------------------------------- */

/* Conv2d(k, stride, pad=(k-1)/2) + bias -> pre-norm output [Hout][Wout][cout_stored], and the GroupNorm sums of
 * the output (groups of 8 channels, models/module.py:101) accumulated into out_gn_sums (may be NULL).
 * weight: [k*k][cin][cout] (tap-major repack of nn.Conv2d's [cout][cin][k][k]); cin = sum of src channels.
 * Replaces nn.Conv2d inside convgnrelu (models/module.py:98-103) and p_conv / m_conv (module.py:118,124,134-136).
 * Supported (cin, cout, k, stride): (4,8,3,1) (8,16,3,1) (16,32,3,1) (32,32,3,1) (32,32,3,2) (32,16,1,1) (32,8,1,1). */
int aarmvs_conv2d(const aarmvs_conv_src* src, int nsrc, int Hin, int Win, int cin, int cout, int ksize, int stride,
                  const float* weight, const float* bias, float* out, int cout_stored, double* out_gn_sums,
                  void* stream);

/* Modulated deformable 3x3 conv, 32 -> 32 channels (models/module.py:133-189, DeformConv2d.forward).
 * x: the layer input (one conv source, usually GN_RELU of the trunk); offmask: [H][W][32] output of the fused
 * p_conv/m_conv (channels 0-8 row offsets, 9-17 column offsets, 18-26 modulation logits, 27-31 unused);
 * weight: [9][32][32] tap-major; out: pre-norm [H][W][32]; out_gn_sums: 4 groups. */
int aarmvs_deform_conv(const aarmvs_conv_src* x, int H, int W, const float* offmask, const float* weight,
                       const float* bias, float* out, double* out_gn_sums, void* stream);

/* The same layer on the tensor cores (gathered A operand staged as fp16 (hi, lo), 8 lanes per pixel and corner line,
 * nine taps accumulated in TMEM).  weight_f16: DEVICE [9 taps][2 K steps][2 chunks][64 rows: 0-31 hi(w[o]), 32-63 lo][8]
 * fp16 (pack.deform_taps_f16); other arguments as aarmvs_deform_conv.  H * W * 128 must stay below 2^32. */
int aarmvs_deform_conv_tc(const aarmvs_conv_src* x, int H, int W, const float* offmask, const void* weight_f16,
                          const float* bias, float* out, double* out_gn_sums, void* stream);

/* IntraViewAAModule tail (models/drmvsnet.py:19-24): GroupNorm+ReLU of the three 1x1 outputs, bilinear x2 / x4
 * upsampling with align_corners=True and channel concat -> feature [H][W][32].  m0: [H][W][16], m1: [H/2][W/2][8],
 * m2: [H/4][W/4][8]; sums/weight/bias as in aarmvs_conv_src (2, 1, 1 groups). */
int aarmvs_feat_assemble(const float* m0, const float* m1, const float* m2, const double* sums0, const double* sums1,
                         const double* sums2, const float* gnw0, const float* gnb0, const float* gnw1,
                         const float* gnb1, const float* gnw2, const float* gnb2, int H, int W, float* feature,
                         void* stream);

/* ---- plane sweep: warp + inter-view adaptive aggregation ------------------------------------------------ */

/* homo_warping_depthwise (models/module.py:6-38) for one source view and one depth: out [H][W][32].
 * cam_host: 12 floats = rot (3x3 row-major) then trans of proj = src_proj @ inverse(ref_proj) (module.py:16-18). */
int aarmvs_warp(const float* src, float* out, int H, int W, const float* cam_host, float depth, int align_corners,
                void* stream);

size_t aarmvs_omega_stats_bytes(int S);    /* S * 3 GroupNorms * (sum, sumsq) doubles PER PLANE; zero before use */
int aarmvs_omega_tail_floats(void);        /* length of the packed omega tail parameter block (floats) */

/* PLANE BATCHING.  Passes A-D have no dependency across depth planes (only the recurrence has), so each call handles
 * `planes` (1..16) consecutive planes d .. d+planes-1 in (tile, view, plane) order: the source-feature lines a tile
 * touches at plane k+1 are almost the ones of plane k and are served from L1/L2 instead of HBM.
 * Layouts: t0 [planes][S][H*W][4], stats [planes][S][3][2] doubles, cost [planes][H*W][32], reweight_out [planes][S][H*W].
 */

/* Pass A: for each of S source views gather at depth_values[d], c = (warp - ref)^2, t0 = conv3x3 32->4 (+bias)
 * (omega.reweight_network.0.0, models/drmvsnet.py:31); writes t0 [S][H*W][4] and GroupNorm-1 sums.
 * src: S feature maps, view v at src + v * view_stride floats; cams: [S][12] device floats;
 * stem_w_host: [9][32][4] tap-major, stem_b_host: [4] (host: baked into the kernel parameter bank). */
int aarmvs_cost_stem(const float* ref, const float* src, size_t view_stride, int H, int W, int S, const float* cams,
                     const float* stem_w_host, const float* stem_b_host, const float* depth_values, int d_host,
                     const int* d_dev, int planes, int align_corners, float* t0, double* stats, void* stream);

/* Pass A with the 3x3 stem conv on the tensor cores (tcgen05 implicit GEMM over the staged (warp - ref)^2 halo tile,
 * split-fp16 operands, fp32 accumulation in TMEM).  Arguments as aarmvs_cost_stem except
 * stem_w_f16: DEVICE [9 taps][4 chunks of 8 input channels][16 rows][8] fp16, rows 0-3 = hi(w[o]), rows 8-11 = lo(w[o])
 * (others zero): per tile the conv is 9 taps x 2 K-steps x (A_hi, A_lo) MMAs with N = 16, taps addressed by descriptor.
 * costv [planes][S][H*W][32] fp32 = (warp - ref)^2; costv_mode 0: unused; 1: WRITTEN here (interior pixels of each tile)
 * so that aarmvs_aggregate can skip its gather -- the gather then runs exactly once per pixel, view and plane.
 * H * W * 128 < 2^31 (the gather recipes hold 32-bit byte offsets): larger images return AARMVS_ERR_ARG.
 * The four lines of a bilinear sample are read at fixed steps from the north-west one (+128, +128 W, +128 W + 128 bytes), a
 * corner outside the image with weight 0: every source map must be followed by (W + 1) * 128 readable bytes holding finite
 * values (the next view's map, or zeroed slack behind the last one -- Engine.feature_buffer); nothing is read before a map. */
int aarmvs_cost_stem_tc_max_slots(void);   /* upper bound of planes * S per launch (shared-memory GroupNorm partial sums) */
int aarmvs_cost_stem_tc(const float* ref, const float* src, size_t view_stride, int H, int W, int S, const float* cams,
                        const void* stem_w_f16, const float* stem_b_host, const float* depth_values, int d_host,
                        const int* d_dev, int planes, int align_corners, float* t0, double* stats, float* costv,
                        int costv_mode, void* stream);

/* costv[k][v][p][:] = (warp_v(p) - ref(p))^2 for the `planes` planes starting at d: homo_warping_depthwise
 * (models/module.py:6-38) + the squared difference (models/drmvsnet.py:311), nothing else.  Used by the backward pass, which
 * re-materialises the per-view cost of one plane (the operand of the omega stem's weight gradient). */
int aarmvs_warp_cost(const float* ref, const float* src, size_t view_stride, int H, int W, int S, const float* cams,
                     const float* depth_values, int d_host, const int* d_dev, int planes, int align_corners, float* costv,
                     void* stream);

/* Pass B (stage 2) / pass C (stage 3): GroupNorm-2 / GroupNorm-3 sums of the omega residual block
 * (models/module.py:256-258) recomputed point-wise from t0.  tail_host: aarmvs_omega_tail_floats() floats. */
int aarmvs_omega_stats(int stage, const float* t0, int H, int W, int S, int planes, const float* tail_host,
                       double* stats, void* stream);

/* Pass D: cost[p][c] = -(sum_v (omega_v(p) + 1) * (warp_v - ref)^2) / S  (models/drmvsnet.py:311-319), the input
 * of the recurrent regulariser.  costv: NULL = gather again (same readable-slack requirement behind the source maps as
 * aarmvs_cost_stem_tc); else pass A's costv_out, read back as a pure stream (ref, src, cams, depth_values are then unused).  reweight_out: optional omega maps (operator seam / tests), may be NULL.
 * cost_planes: write the result as fp16 (hi, lo) operand planes [planes][2][4][H][W][8] (see aarmvs_plane_src) instead of
 * fp32 `cost`, which may then be NULL. */
int aarmvs_aggregate(const float* ref, const float* src, size_t view_stride, int H, int W, int S, const float* cams,
                     const float* tail_host, const float* depth_values, int d_host, const int* d_dev, int planes,
                     int align_corners, const float* t0, const double* stats, const float* costv, float* cost,
                     void* cost_planes, float* reweight_out, void* stream);

/* ---- recurrent regularisation (models/drmvsnet.py:119-167) --------------------------------------------- */

/* ConvLSTMCell.forward (models/module.py:76-92).  Input = concat(src[0..nsrc-1], h_{d-1}); cin counts all of
 * them.  weight_packed: [9][cin][4*hidden] with output channel order [hidden channel][i,f,o,g];
 * bias_packed: [4*hidden] same order.  h_pair: two [H*W][hidden] halves; c_state: [H*W][hidden], in place.
 * Plane 0 uses h = c = 0 without reading the buffers (module.py:94-96, drmvsnet.py:133-134).
 * Supported (cin, hidden): (48,16) (32,16) (40,8). */
int aarmvs_convlstm_cell(const aarmvs_conv_src* src, int nsrc, int H, int W, int cin, int hidden,
                         const float* weight_packed, const float* bias_packed, float* h_pair, float* c_state,
                         int d_host, const int* d_dev, void* stream);

/* Operand planes: how the tensor-core cell kernel receives every input.  A C-channel fp32 map x [H][W][C] is stored
 * as two fp16 planes with x ~= hi + lo (hi = x with the low 13 mantissa bits cleared, lo = fp16(x - hi), 2^-22 relative):
 *   [pp ping-pong halves][2 (hi, lo)][C/8 chunks][H][W][8 channels] fp16
 * so that one 16-byte element is 8 channels of one pixel = one row of a K-major tcgen05 core matrix and a TMA box of
 * (10 x 18 pixels x chunks) IS the staged halo tile of the implicit GEMM.  Producers write this layout directly
 * (aarmvs_aggregate's cost_planes, the cell's own h / pooled outputs, aarmvs_split_planes). */
typedef struct aarmvs_plane_src {
    const void* planes;        /* device pointer, 16-byte aligned */
    int channels;              /* multiple of 8 */
    int pp;                    /* 1 = single buffer, 2 = ping-pong pair */
    int pp_phase;              /* pp == 2: 0 = half written by the current plane (d & 1), 1 = previous plane's half */
} aarmvs_plane_src;

/* fp32 [H][W][channels] (mode PLAIN or GN_RELU: the producer's GroupNorm + ReLU is applied first, models/module.py:284-286)
 * -> operand planes [2][channels/8][H][W][8]. */
int aarmvs_split_planes(const aarmvs_conv_src* src, int H, int W, void* planes, void* stream);

/* Floats of the cell-state buffer of aarmvs_convlstm_cell_tc: c lives in a tile-blocked layout private to the kernel,
 * [tiles of 16 rows x 8 columns, row-major over ceil(H/16) x ceil(W/8)][4 quarters of the hidden channels]
 * [128 pixels of the tile, row-major][hidden/4 channels]  (every access of the epilogue is a fully used 128-byte line). */
size_t aarmvs_cell_state_floats(int H, int W, int hidden);

/* The same cell on the tensor cores (tcgen05.mma kind::f16, fp32 accumulation in TMEM), fed by TMA from operand planes.
 * Each K step issues hi*hi + lo*hi + hi*lo, which keeps ~2^-21 relative accuracy per product (single-pass TF32 / BF16
 * fails the parity bar).  Input = concat(src[0..nsrc-1], h_{d-1}); cin counts all of them.
 * weight_f16: [K steps][2 chunks of 8 input channels][2 (hi, lo)][4*hidden][8] fp16, output-channel order
 * [hidden channel][i,f,o,g]: per tap one step per 16 input channels; with cin = 40 the fifth chunk of the tap pairs
 * (0,1) (2,3) (4,5) (6,7) (8,-) shares one step each, after the 18 regular ones (23 steps; pack.lstm_gates_f16_steps);
 * bias_packed as aarmvs_convlstm_cell.
 * h_planes: ping-pong operand planes [2][2][hidden/8][H][W][8]; plane d reads half (d+1)&1 as h_{d-1} (not at d = 0,
 * where h = c = 0) and writes half d&1.  c_blocked: aarmvs_cell_state_floats() floats, in place.
 * h_f32_pair: optional fp32 copy of h, two [H*W][hidden] halves (half d&1 written) for the deconvolution / head kernels.
 * pooled_planes: optional operand planes [2][hidden/8][H/2][W/2][8] of max_pool2d(h, 2) (models/drmvsnet.py:148,152;
 * hidden == 16 only).
 * gates_out: optional (training forward); *gates_out (device memory, like aarmvs_snapshot_state's state_pp, so one captured
 * graph serves every sample) is the base of a history [num planes][H*W][4*hidden] fp32 of the ACTIVATED gates in channel blocks
 * sigmoid(i) | sigmoid(f) | sigmoid(o) | tanh(g) (module.py:83-88), record d written by plane d: the backward pass reads the
 * gates exactly as the forward used them instead of recomputing the cell's convolution.
 * Supported (cin, hidden): (48,16) (32,16) (40,8). */
int aarmvs_convlstm_cell_tc(const aarmvs_plane_src* src, int nsrc, int H, int W, int cin, int hidden,
                            const void* weight_f16, const float* bias_packed, void* h_planes, float* c_blocked,
                            float* h_f32_pair, void* pooled_planes, float* const* gates_out, int d_host, const int* d_dev, void* stream);

/* deConvGnReLU's ConvTranspose2d(16,16,k3,s2,p1,op1) (models/module.py:281): in [Hin][Win][16] (one conv source,
 * typically a ping-pong h) -> pre-norm out [2Hin][2Win][16] + GroupNorm sums (2 groups).
 * weight_host: [3][3][16 cin][16 cout] repack of ConvTranspose2d's [cin][cout][kh][kw], bias_host: [16] -- HOST
 * memory: both are baked into the kernel parameter (constant) bank so the FMAs need no weight loads. */
int aarmvs_deconv2x(const aarmvs_conv_src* in, int Hin, int Win, const float* weight_host, const float* bias_host,
                    float* out, double* out_gn_sums, int d_host, const int* d_dev, void* stream);

/* The same layer on the tensor cores, fed by TMA from the operand planes of h (16 channels) and run twice per plane
 * because GroupNorm (models/module.py:284) needs the sums of the whole output map:
 *   out_planes == NULL: transposed conv -> gn_sums [2 groups][sum, sumsq] += (nothing is stored; zero gn_sums before);
 *   out_planes != NULL: transposed conv again -> GroupNorm(gn_sums, gn_weight, gn_bias) + ReLU -> fp16 split ->
 *                       operand planes [2][2][2Hin][2Win][8] of the next cell.
 * weight_f16: DEVICE [9 (class, tap) pairs][2 chunks][32 rows: 0-15 hi(W[.][co]), 16-31 lo][8] fp16 in the pair order
 * W[1][1], W[1][2], W[1][0], W[2][1], W[0][1], W[2][2], W[2][0], W[0][2], W[0][0] of ConvTranspose2d's [cin][cout][ky][kx];
 * bias_host: [16] host memory. */
int aarmvs_deconv2x_tc(const aarmvs_plane_src* in, int Hin, int Win, const void* weight_f16, const float* bias_host,
                       double* gn_sums, const float* gn_weight, const float* gn_bias, void* out_planes, int d_host,
                       const int* d_dev, void* stream);

/* conv_0 (3x3, 8 -> 1; models/drmvsnet.py:165) fused with the winner-take-all update (drmvsnet.py:324-339):
 * p = exp(cost); if (max_prob < p) {max_prob = p; depth = depth_values[d]; index = d}; exp_sum += p.
 * At d == num_depth - 1 also writes conf = max_prob / exp_sum.  At d == 0 the state maps are treated as zero.
 * weight_host: [9][8] tap-major, bias_host: [1] (host memory: baked into the kernel parameters).
 * cost_volume: optional [num_depth][H*W] store of the regularised cost (training phase / tests), may be NULL. */
int aarmvs_head_wta(const aarmvs_conv_src* h4, int H, int W, const float* weight_host, const float* bias_host,
                    const float* depth_values, int num_depth, int d_host, const int* d_dev, float* max_prob,
                    float* depth_map, float* exp_sum, int* index_map, float* conf, float* cost_volume, void* stream);

/* The same head on the tensor cores, fed by TMA from h4's ping-pong operand planes [2][2][1][H][W][8] (half d & 1 is
 * read): two taps per K = 16 MMA.  weight_f16: DEVICE [5 pairs][2 taps][16 rows][8] fp16 with row 0 = hi(w), row 8 = lo(w)
 * for the tap pairs (0,1) (2,3) (4,5) (6,7) (8,-) of tap = ky * 3 + kx; other arguments as aarmvs_head_wta.
 * Optional fused training loss (mvsnet_cls_loss, models/drmvsnet.py:347-381, without the [D][H][W] probability volume):
 * gt_index [H*W] = hypothesis nearest to the ground truth (0 where gt_mask is 0), gt_mask [H*W], gt_cost [3][H*W] state (cost
 * at the ground-truth plane | running max | rescaled exp-sum); at the last plane
 * *loss_sum += sum_p gt_mask[p] * (log(sum_d exp(cost_d[p])) - cost_gt[p]), the log-sum-exp evaluated as an online softmax, so
 * the loss stays finite where the reference's F.softmax does.  All four NULL: no loss.
 * Range of the winner-take-all part: drmvsnet.py:324 takes exp(cost) without subtracting a maximum, so like the reference's
 * test phase max_prob / exp_sum overflow for a regularised cost above ~88; the shipped weights produce costs in [-41, -17]
 * (SURVEY.md §7).  The fp16 (hi, lo) operand split of the tensor-core kernels assumes |x| < 65 504. */
int aarmvs_head_wta_tc(const void* h4_planes, int H, int W, const void* weight_f16, const float* bias_host,
                       const float* depth_values, int num_depth, int d_host, const int* d_dev, float* max_prob,
                       float* depth_map, float* exp_sum, int* index_map, float* conf, float* cost_volume,
                       const int* gt_index, const float* gt_mask, float* gt_cost, double* loss_sum, void* stream);

/* End of a plane: ++*d_dev and zero `nzero` doubles of GroupNorm sums for the next plane. */
int aarmvs_next_plane(int* d_dev, double* sums, int nzero, void* stream);

/* ---- depth-map fusion filter (the step after the path; fusion.py:71-133, 179-225) ------------------------ */

/* Geometric-consistency filter of one reference view against S <= 10 source views (reproject with the reference depth,
 * sample the source depth like cv2.remap INTER_LINEAR, project back, nine nested thresholds dist < i/4 px and
 * |d_rep - d| / d < i/1300, i = 2..10; geo_mask = OR_i (views passing i >= i) for i < S + 1; averaged depth
 * (sum of the consistent reprojected depths + d) / (count + 1)).
 * mats: DEVICE [S][aarmvs_fuse_view_mats() = 60] float64 per source view: inv(K_ref)[9], (E_src inv(E_ref))[3x4],
 * K_src[9], inv(K_src)[9], (E_ref inv(E_src))[3x4], K_ref[9], each prepared in float32 as numpy does (fusion.py:80-105).
 * depth_src: [S][H*W].  final_mask = geo_mask && conf_ref > photo_threshold (may be NULL with conf_ref), geo_sum may be NULL. */
int aarmvs_fuse_view_mats(void);
int aarmvs_fuse_view(const float* depth_ref, const float* conf_ref, const float* depth_src, int S, int H, int W,
                     const double* mats, float photo_threshold, unsigned char* geo_mask, unsigned char* final_mask,
                     float* depth_avg, int* geo_sum, void* stream);

/* ---- tensor-core building blocks ------------------------------------------------------------------------- */

/* Training tail, first step of the backward pass (models/drmvsnet.py:289-291 softmax over the D regularised costs, :347-381
 * mvsnet_cls_loss): d loss / d cost for every plane from the stored regularised costs,
 *   dcost[d][p] = scale * mask[p] * (softmax_d(cost[:, p]) - [d == gt_index[p]]),  scale = 1 / ((sum_p mask[p] + 1e-6) * B).
 * cost, dcost: [D][P] fp32 (plane major, one sample); gt_index: [P] int32 = round(mask * argmin_d |depth_values[d] - depth_gt|);
 * mask: [P] fp32. */
int aarmvs_loss_grad(const float* cost, const int* gt_index, const float* mask, float scale, int D, size_t P, float* dcost,
                     void* stream);

/* Known-answer test of the tcgen05 path shared by the ConvLSTM kernels (un-swizzled K-major shared-memory operands,
 * 8-row groups `group_rows` rows apart, start row `shift`, TMEM accumulation, commit -> mbarrier, tcgen05.ld):
 *   d[m][n] = sum_k a[shift + (m / 8) * group_rows + m % 8][k] * b[n][k],  m < 128, n < 64, fp16 inputs, fp32 out.
 * a_f16: [a_rows][K] row-major fp16, b_f16: [64][K]; K multiple of 16, <= 256. */
int aarmvs_tc_selftest(const void* a_f16, int a_rows, const void* b_f16, float* d, int K, int group_rows, int shift,
                       void* stream);

/* Debug micro-benchmark: clock64 cycles for reps*9 back-to-back tcgen05.mma (M=128, N, K=16, fp16, un-swizzled operands
 * with 8-row groups `group_rows` rows apart, start row advancing by a_shift_rows per instruction) on `grid` CTAs. */
int aarmvs_tc_mma_bench(int N, int group_rows, int reps, int a_shift_rows, int grid, long long* out, void* stream);

/* ---- backward pass (BASELINE config 5: train.py:288-307 loss.backward() through models/drmvsnet.py:272-295) ---------
 * fp32 NHWC primitives; the host side (aarmvsnet_b200/backward.py) walks the depth planes in reverse, recomputes each
 * plane's intermediates from the stored recurrent state and chains these calls.  Gradient buffers marked "+=" are
 * accumulated into (zero them first); nothing here allocates. */

/* One channel segment of a concatenated (inputs) or split (outputs) NHWC operand: [H][W][channels] fp32. */
typedef struct aarmvs_seg {
    float* ptr;
    int channels;
    int accumulate;            /* outputs only: 0 = store, 1 = add to what is there */
} aarmvs_seg;

/* Direct convolution over the channel concat of `in` (replaces torch.cat, module.py:80), written to / added into the channel
 * split `out`:
 *   transposed == 0:  out(Y,X)[co] = bias[co] + sum in(stride*Y - pad + ky, stride*X - pad + kx)[ci] * weight[ky*k+kx][ci][co]
 *   transposed == 1:  out(Y,X)[co] = bias[co] + sum in((Y + pad - ky) / stride, (X + pad - kx) / stride)[ci] * weight[..][ci][co]
 *                     over the taps where the divisions are exact;  pad = (k - 1) / 2, zero outside the input.
 * Used for: gate pre-activations recomputed per plane (module.py:82), dX of every Conv2d (transposed, weight =
 * W[co][ci][ky][kx] read as [tap][co][ci]), ConvTranspose2d forward (module.py:281: transposed, stride 2) and its dX
 * (plain, stride 2), the 1x1 layers.  weight: [k*k][cin][cout]; bias: [cout] or NULL; k in {1, 3}; stride in {1, 2}. */
int aarmvs_conv_nhwc(const aarmvs_seg* in, int nin, int Hin, int Win, const aarmvs_seg* out, int nout, int Hout, int Wout,
                     int cin, int cout, int ksize, int stride, int transposed, const float* weight, const float* bias,
                     void* stream);

/* dweight[ky*k+kx][ci][co] += sum_{Y,X} x(stride*Y - pad + ky, stride*X - pad + kx)[ci] * dy(Y,X)[co];  dbias[co] += sum dy.
 * x: channel concat [Hx][Wx][cx]; dy: [Hy][Wy][cy], cy <= 64.  For a Conv2d x is its input and dy its output gradient; for
 * the ConvTranspose2d x is the OUTPUT gradient and dy the input (the result is then indexed [tap][co][ci]). */
int aarmvs_conv_bwd_weight(const aarmvs_seg* x, int nx, int Hx, int Wx, const float* dy, int Hy, int Wy, int cx, int cy,
                           int ksize, int stride, float* dweight, float* dbias, void* stream);

/* GroupNorm of x [P][C] (groups of C / groups channels over all P pixels; module.py:101,284): sums [groups][2] += (sum, sum
 * of squares);  y = [ReLU](gamma * (x - mean) * rstd + beta).  C divides 256 (4, 8, 16, 32, 64: every layer of the path). */
int aarmvs_gn_sums(const float* x, size_t P, int C, int groups, double* sums, void* stream);
int aarmvs_gn_relu_apply(const float* x, size_t P, int C, int groups, const double* sums, const float* gamma, const float* beta,
                         int relu, float* y, void* stream);
/* Backward of the same: stage 1 accumulates chan_sums [C][2] (zero before) = (sum_p dz * xhat, sum_p dz), dz = dy * [y > 0];
 * stage 2 writes dx (may alias dy), adds the affine gradients to dgamma / dbeta and, if dx_chan_sum != NULL, sum_p dx to it
 * (the bias gradient of the convolution that produced x).  C divides 256. */
int aarmvs_gn_relu_bwd(int stage, const float* x, const float* dy, size_t P, int C, int groups, const double* sums,
                       const float* gamma, const float* beta, int relu, double* chan_sums, float* dx, float* dgamma,
                       float* dbeta, float* dx_chan_sum, void* stream);

/* ConvLSTMCell backward, point-wise part (module.py:83-90).  gates [P][4*hidden]: the conv output incl. bias, channel blocks
 * i | f | o | g; c_prev (NULL at plane 0) / c_new [P][hidden]; dh = dL/dh_d; dc: in dL/dc_d (from plane d+1), out dL/dc_{d-1};
 * dgates [P][4*hidden] = dL/d(conv output). */
int aarmvs_lstm_gates_bwd(const float* gates, const float* c_prev, const float* c_new, const float* dh, float* dc, float* dgates,
                          size_t P, int hidden, void* stream);
/* The same from the ACTIVATED gates the training forward stored (aarmvs_convlstm_cell_tc: gates_out): *gates_history (device
 * memory) is the base of [planes][P][4*hidden]; record d = *d_dev if d_dev is non-null (one captured graph serves every plane
 * and sample), else d_host.
 * max_bits: optional; receives max |dgates| as the bit pattern of a float (atomicMax on a zeroed word), the scale of
 * aarmvs_split_planes_scaled.  dbias: optional [4*hidden], += sum_p dgates (the cell conv's bias gradient). */
int aarmvs_lstm_gates_bwd_act(const float* const* gates_history, int d_host, const int* d_dev, const float* c_prev, const float* c_new,
                              const float* dh, float* dc, float* dgates, size_t P, int hidden, unsigned* max_bits, float* dbias,
                              void* stream);

/* The cells' data gradient dX = dG (*) W^T on the tensor cores (csrc/conv_dx_tc.cu): the forward cell's TMA-fed tcgen05
 * implicit GEMM with the gate gradients as operand planes.
 * aarmvs_split_planes_scaled: fp32 x [H][W][channels] -> operand planes [2][channels/8][H][W][8] of 2^k x, k chosen so that
 * the largest element (*max_bits, see above) lands in [128, 256): gradients are far below fp16's normal range; writes 2^-k
 * to *inv_scale (device).
 * aarmvs_conv_dx_tc: out[Y][X][co] = 2^-k sum_{ky,kx,ci} planes[Y-1+ky][X-1+kx][ci] * w[ky*3+kx][ci][co] scattered to
 * `out` like aarmvs_conv_nhwc (channels multiples of 4, stored or accumulated).  weight_f16: DEVICE
 * [9 * cin/16 K steps][2 chunks][2 (hi, lo)][NG][8] fp16 with NG = 48 or 32 output rows (zero rows beyond cout) -- for the
 * data gradient the caller packs the spatially flipped, transposed cell weights (pack.conv_f16_steps).
 * Supported (cin, cout): (64,48) (64,32) (32,40). */
int aarmvs_split_planes_scaled(const float* x, int H, int W, int channels, const unsigned* max_bits, void* planes,
                               float* inv_scale, void* stream);
/* The cells' weight gradient dW[tap][ci][co] += 2^-k sum_p X[p + tap][ci] * dG[p][co] on the tensor cores
 * (csrc/conv_dw_tc.cu): a K = pixels GEMM over the same operand planes read as MN-major operands, accumulated in TMEM over
 * all tiles of a CTA.  x_planes: [2][cx/8][H][W][8], the channel concat [x | h_{d-1}] of the cell's inputs
 * (aarmvs_split_planes_cat: up to three fp32 [H][W][ch] maps, channels multiples of 8, unscaled);
 * dgate_planes / inv_scale: aarmvs_split_planes_scaled's.  dweight: [9][cx][cy] fp32, accumulated with atomics.
 * Supported (cx, cy): (48,64) (32,64) (40,32).  The bias gradient is aarmvs_lstm_gates_bwd_act's dbias. */
int aarmvs_split_planes_cat(const aarmvs_seg* x, int nx, int H, int W, void* planes, void* stream);
int aarmvs_conv_dw_tc(const void* x_planes, const void* dgate_planes, int H, int W, int cx, int cy, const float* inv_scale,
                      float* dweight, void* stream);
int aarmvs_conv_dx_tc(const void* dgate_planes, int H, int W, int cin, const aarmvs_seg* out, int nout, int cout,
                      const void* weight_f16, const float* inv_scale, void* stream);

/* MaxPool2d(2, 2) (drmvsnet.py:148,152) and its backward (dx += dout at the first maximum of each window). */
int aarmvs_maxpool2(const float* x, int H, int W, int C, float* out, void* stream);
int aarmvs_maxpool2_bwd(const float* x, const float* dout, int H, int W, int C, float* dx, void* stream);

/* Aggregation backward (drmvsnet.py:311-319): cost = -(1/S) sum_v (1 + omega_v) c_v  ->
 *   dcostv[v][p][:] = -(1 + reweight[v][p]) / S * dcost[p][:],   dreweight[v][p] = -(1/S) sum_ch costv[v][p][ch] * dcost[p][ch]. */
int aarmvs_aggregate_bwd(const float* dcost, const float* costv, const float* reweight, int S, size_t P, float* dcostv,
                         float* dreweight, void* stream);

/* Backward of omega's point-wise tail (drmvsnet.py:31-34 after the stem conv), four passes because each of its three
 * GroupNorms needs two sums over the whole map: stage 3, 2, 1 accumulate bstats[v][stage-1][2] (zero before stage 3) and the
 * parameter gradients dtail[aarmvs_omega_tail_floats()] (fp64, +=, OmegaTail order); stage 0 writes dt0 [S][P][4] = dL/d(stem
 * conv output).  t0, stats, tail_host as in aarmvs_omega_stats (one plane); dreweight from aarmvs_aggregate_bwd. */
int aarmvs_omega_bwd(int stage, const float* t0, const float* dreweight, int H, int W, int S, const float* tail_host,
                     const double* stats, double* bstats, double* dtail, float* dt0, void* stream);

/* Warp backward (module.py:6-38; the grid is built under no_grad, so only the features receive gradients): with
 * c_v = (warp_v - ref)^2:  g = 2 (warp_v - ref) * dcostv[v];  dref[p] -= sum_v g;  dsrc[v][tap] += w_tap * g for the four
 * bilinear taps (red.global.add.v4.f32).  cams: device [S][12]; the plane's depth is depth_values[d] with d = *d_dev if d_dev
 * is non-null (one captured graph serves every plane), else d_host; dsrc: S maps `view_stride` floats apart. */
int aarmvs_warp_bwd(const float* ref, const float* src, size_t view_stride, int H, int W, int S, const float* cams,
                    const float* depth_values, int d_host, const int* d_dev, int align_corners, const float* dcostv, float* dref,
                    float* dsrc, void* stream);

/* IntraViewAAModule tail backward (drmvsnet.py:19-24): dfeature [H][W][32] -> gradients w.r.t. the GroupNorm+ReLU'd 1x1
 * outputs: dm0 [H][W][16] (stored), dm1 [H/2][W/2][8] and dm2 [H/4][W/4][8] (+=, atomics: zero before). */
int aarmvs_feat_assemble_bwd(const float* dfeature, int H, int W, float* dm0, float* dm1, float* dm2, void* stream);

/* DeformConv2d backward (module.py:133-189), with the 9x32 -> 32 contraction done by aarmvs_conv_nhwc / _bwd_weight (k = 1):
 * aarmvs_deform_sample: x [H][W][32] (already GroupNorm+ReLU'd), offmask as in aarmvs_deform_conv ->
 *     samp [P][9][32] (bilinear samples before the modulation) and samp_m = samp * sigmoid(m)  (the contraction's input);
 * aarmvs_deform_bwd: dsamp_m [P][9][32] -> dx [H][W][32] (+=, atomics) and doffmask [P][32] (row / column offsets through the
 *     bilinear weights -- floor() is detached, module.py:150 -- and modulation logits; channels 27-31 = 0). */
int aarmvs_deform_sample(const float* x, int H, int W, const float* offmask, float* samp, float* samp_m, void* stream);
int aarmvs_deform_bwd(const float* x, int H, int W, const float* offmask, const float* samp, const float* dsamp_m, float* dx,
                      float* doffmask, void* stream);

/* Training forward: after plane d, copy the recurrent state of the five cells -- h_i (the fp32 copies written by
 * aarmvs_convlstm_cell_tc, two [P_i][Ch_i] halves each; half d & 1 is read) and c_i (tile-blocked, aarmvs_cell_state_floats) --
 * into record r of the state history [D + 1][record_floats] as NHWC fp32 [h0 | c0 | h1 | c1 | h2 | c2 | h3 | c3 | h4 | c4]
 * (hidden 16, 16, 16, 16, 8 at scales 1, 2, 4, 2, 1: 66 floats per full-resolution pixel).  r = *d_dev if d_dev is non-null
 * (the plane counter AFTER aarmvs_next_plane, i.e. d + 1), else d_host; *state_pp (device memory) is the history's base, so one
 * captured graph serves every sample.  h_f32_pairs / c_blocked: HOST arrays of five device pointers. */
int aarmvs_snapshot_state(const float* const* h_f32_pairs, const float* const* c_blocked, int H, int W, float* const* state_pp,
                          size_t record_floats, int d_host, const int* d_dev, void* stream);

/* Adam step (train.py:180,307; torch.optim.Adam semantics, no amsgrad) over one flat fp32 buffer holding every parameter:
 * param, grad, exp_avg, exp_avg_sq [n]; step = 1 for the first update. */
int aarmvs_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, int step, void* stream);

/* ---- debug hooks: only in libraries built with -DAARMVS_DEBUG (AARMVS_NVCC_EXTRA=-DAARMVS_DEBUG python -m
 * aarmvsnet_b200.build --force); the product library has no process-global state and does not export them. */
#ifdef AARMVS_DEBUG
/* Debug ablation switches for aarmvs_cost_stem_tc (results become wrong; timing experiments only): 2 skip the sample-position math, 4 skip the MMAs.  0 = normal. */
void aarmvs_debug_flags(int flags);
/* Debug: worker group 0 of CTA 0 of subsequent aarmvs_cost_stem_tc launches records clock64 stamps per fill ([64][8]
 * uint64 device buffer; NULL switches off). */
void aarmvs_debug_stem_profile_buffer(void* buf);

/* Debug aid: when set to a device buffer of
 * 64 x 12 uint64, CTA 0 of subsequent aarmvs_convlstm_cell_tc launches records clock64 stamps of its loader / MMA /
 * epilogue roles per tile.  Pass NULL to switch off. */
void aarmvs_debug_profile_buffer(void* buf);
/* Debug ablation switches for aarmvs_convlstm_cell_tc (results become wrong): 1 no TMA loads,
 * 4 no global I/O in the epilogue, 8 no MMAs, 16 no gate transcendentals, 32 no max-pool output. */
void aarmvs_debug_cell_flags(int flags);

#endif

#ifdef __cplusplus
}
#endif
#endif
