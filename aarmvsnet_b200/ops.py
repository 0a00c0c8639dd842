"""Operator seams with the reference's call signatures (NCHW tensors in and out), each running the CUDA kernels.

These are the individual Python call sites a maintainer could hook: `homo_warping_depthwise`
(models/module.py:6), `ConvLSTMCell.forward` (module.py:76), `FeatNet.forward` (drmvsnet.py:56), the per-plane
aggregation (drmvsnet.py:307-318).  The fused forward in engine.py does not go through them; they exist so the
kernels can be checked one at a time against the oracle.  Layout conversion NCHW <-> NHWC happens here.
"""
import ctypes

import torch

from . import capi
from .engine import Engine, _psrc, _psrc_array, _ptr, _src, _src_array


def _nhwc(x):          # [C,H,W] -> [H*W, C] contiguous
    return x.permute(1, 2, 0).reshape(-1, x.shape[0]).contiguous()


def _nchw(x, h, w):    # [H*W, C] -> [C,H,W]
    return x.view(h, w, -1).permute(2, 0, 1).contiguous()


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _need_cuda(t):
    if t.device.type != "cuda":
        raise RuntimeError("aarmvsnet_b200 has no CPU path: tensors must live on a CUDA device")


def homo_warping_depthwise(src_fea, src_proj, ref_proj, depth_value, align_corners=False):
    """models/module.py:6-38.  src_fea [B,32,H,W], src_proj/ref_proj [B,4,4], depth_value [B] -> [B,32,H,W]."""
    _need_cuda(src_fea)
    lib = capi.load()
    B, C, H, W = src_fea.shape
    if C != 32:
        raise RuntimeError("the warp kernel is specialised for the 32-channel feature maps of FeatNet")
    out = torch.empty_like(src_fea)
    for b in range(B):
        cam = Engine.relative_cameras(torch.stack([ref_proj[b], src_proj[b]]))[0]
        src = _nhwc(src_fea[b].float())
        dst = torch.empty_like(src)
        capi.check(lib.aarmvs_warp(_ptr(src), _ptr(dst), H, W, capi.host_floats(cam.tolist()), float(depth_value[b]),
                                   int(align_corners), _stream(src.device)), "aarmvs_warp")
        out[b] = _nchw(dst, H, W)
    return out


def feat_net(model, img):
    """FeatNet.forward (models/drmvsnet.py:56-63).  img [B,3,H,W] -> [B,32,H,W]."""
    _need_cuda(img)
    eng = model._engine
    B, _, H, W = img.shape
    w = eng._weights(img.device)
    fs = eng._feat_scratch(H, W, img.device)
    feat = torch.empty(H * W, 32, device=img.device)
    out = torch.empty(B, 32, H, W, device=img.device)
    for b in range(B):
        eng.feature_map(img[b].float().contiguous(), feat, fs, w, _stream(img.device))
        out[b] = _nchw(feat, H, W)
    return out


def aggregate_plane(model, ref_fea, src_feas, ref_proj, src_projs, depth_value, stream_costv=True):
    """models/drmvsnet.py:307-318 for one sample: ref_fea [32,H,W], src_feas list of [32,H,W], projections [4,4].
    Returns (volume_variance [32,H,W], reweight maps [S,H,W])."""
    _need_cuda(ref_fea)
    eng = model._engine
    lib, dev = eng.lib, ref_fea.device
    w = eng._weights(dev)
    _, H, W = ref_fea.shape
    S, P = len(src_feas), H * W
    feat = Engine.feature_buffer(S + 1, H, W, dev)               # with the readable slack the gather kernels require behind it
    feat.copy_(torch.stack([_nhwc(f.float()) for f in [ref_fea] + list(src_feas)]))
    cams = Engine.relative_cameras(torch.stack([ref_proj] + list(src_projs))).to(dev)
    depths = torch.tensor([float(depth_value)], device=dev)
    t0 = torch.empty(S, P, 4, device=dev)
    sums = torch.zeros(S * 6, device=dev, dtype=torch.float64)
    cost = torch.empty(P, 32, device=dev)
    rew = torch.empty(S, P, device=dev)
    st, ac = _stream(dev), int(eng.align_corners)
    cv = torch.empty(S, P, 32, device=dev) if (eng.cell_math == "tc" and stream_costv) else None
    costv = _ptr(cv) if cv is not None else None
    eng.cost_stem(_ptr(feat), _ptr(feat, P * 32), P * 32, H, W, S, _ptr(cams), w, _ptr(depths), 0, None, 1, _ptr(t0), _ptr(sums), st, costv)
    for stage in (2, 3):
        capi.check(lib.aarmvs_omega_stats(stage, _ptr(t0), H, W, S, 1, w["tail_host"], _ptr(sums), st), "aarmvs_omega_stats")
    capi.check(lib.aarmvs_aggregate(_ptr(feat), _ptr(feat, P * 32), P * 32, H, W, S, _ptr(cams), w["tail_host"],
                                    _ptr(depths), 0, None, 1, ac, _ptr(t0), _ptr(sums), costv, _ptr(cost), None, _ptr(rew), st),
               "aarmvs_aggregate")
    return -_nchw(cost, H, W), rew.view(S, H, W)


def _block_state(c, H, W):
    """[H*W, CH] NHWC cell state -> the tile-blocked layout of aarmvs_convlstm_cell_tc (include/aarmvs.h)."""
    ch = c.shape[1]
    ty, tx = (H + 15) // 16, (W + 7) // 8
    full = c.new_zeros(ty * 16, tx * 8, ch)
    full[:H, :W] = c.view(H, W, ch)
    return full.view(ty, 16, tx, 8, 4, ch // 4).permute(0, 2, 4, 1, 3, 5).contiguous().view(-1)


def _unblock_state(cb, H, W, ch):
    ty, tx = (H + 15) // 16, (W + 7) // 8
    full = cb.view(ty, tx, 4, 16, 8, ch // 4).permute(0, 3, 1, 4, 2, 5).reshape(ty * 16, tx * 8, ch)
    return full[:H, :W].reshape(H * W, ch).contiguous()


def convlstm_cell(model, cell_index, x, h, c, math=None):
    """ConvLSTMCell.forward (models/module.py:76-92) of cost_regularization.cell_list[cell_index].
    x [Cx,H,W], h, c [Ch,H,W] -> (h_next, c_next)."""
    _need_cuda(x)
    eng = model._engine
    lib, dev = eng.lib, x.device
    w = eng._weights(dev)
    cx, H, W = x.shape
    ch = h.shape[0]
    xs = _nhwc(x.float())
    h_pair = torch.empty(2, H * W, ch, device=dev)
    h_pair[0] = _nhwc(h.float())                       # plane d=1 reads half 0 and writes half 1
    cs = _nhwc(c.float())
    st = _stream(dev)
    if (math or eng.cell_math) == "tc":
        # the tensor-core kernel reads fp16 (hi, lo) operand planes and keeps c tile-blocked: convert at the seam
        xp = torch.empty(2, cx // 8, H * W, 8, device=dev, dtype=torch.float16)
        hp = torch.zeros(2, 2, (ch + 7) // 8, H * W, 8, device=dev, dtype=torch.float16)
        eng.split_planes(_src(xs, cx), H, W, xp, st)
        if ch % 8 == 0:
            eng.split_planes(_src(h_pair[0], ch), H, W, hp[0], st)
        cb = _block_state(cs, H, W)
        assert cb.numel() == lib.aarmvs_cell_state_floats(H, W, ch)
        capi.check(lib.aarmvs_convlstm_cell_tc(_psrc_array(_psrc(xp, cx)), 1, H, W, cx + ch, ch, _ptr(w[f"cell{cell_index}.wtc"]),
                                               _ptr(w[f"cell{cell_index}.b"]), _ptr(hp), _ptr(cb), _ptr(h_pair), None, None, 1, None, st),
                   "aarmvs_convlstm_cell_tc")
        cs = _unblock_state(cb, H, W, ch)
    else:
        capi.check(lib.aarmvs_convlstm_cell(_src_array(_src(xs, cx)), 1, H, W, cx + ch, ch, _ptr(w[f"cell{cell_index}.w"]),
                                            _ptr(w[f"cell{cell_index}.b"]), _ptr(h_pair), _ptr(cs), 1, None, st),
                   "aarmvs_convlstm_cell")
    return _nchw(h_pair[1], H, W), _nchw(cs, H, W)


def deconv_gn_relu(model, index, x):
    """deConvGnReLU.forward (models/module.py:269-287) of cost_regularization.deconv_{index} on the tensor cores.
    x [16,H,W] -> [16,2H,2W] (the kernel's fp16 (hi, lo) operand planes summed back to fp32)."""
    _need_cuda(x)
    eng = model._engine
    lib, dev = eng.lib, x.device
    w = eng._weights(dev)
    _, H, W = x.shape
    st = _stream(dev)
    xp = torch.empty(2, 2, H * W, 8, device=dev, dtype=torch.float16)
    eng.split_planes(_src(_nhwc(x.float()), 16), H, W, xp, st)
    out = torch.empty(2, 2, 4 * H * W, 8, device=dev, dtype=torch.float16)
    sums = torch.zeros(4, device=dev, dtype=torch.float64)
    g = w[f"deconv{index}.gn"]
    src = _psrc(xp, 16)
    for o in (None, out):
        capi.check(lib.aarmvs_deconv2x_tc(ctypes.byref(src), H, W, _ptr(w[f"deconv{index}.wtc"]), w[f"deconv{index}.b_host"],
                                          _ptr(sums), _ptr(g[0]), _ptr(g[1]), _ptr(o) if o is not None else None, 1, None, st),
                   "aarmvs_deconv2x_tc")
    y = out.float().sum(0)                                   # hi + lo: [2 chunks][4HW][8]
    return y.permute(1, 0, 2).reshape(2 * H, 2 * W, 16).permute(2, 0, 1).contiguous()


def deform_conv(model, level, x, math=None):
    """DeformConv2d.forward (models/module.py:133-189) of feature.intraAA.deformconv{level}[0].  x [32,H,W] -> [32,H,W].
    math: "tc" (tcgen05, gathered split-fp16 operand) or "fma" (fp32 CUDA cores); default = the engine's cell_math."""
    _need_cuda(x)
    eng = model._engine
    lib, dev = eng.lib, x.device
    w = eng._weights(dev)
    _, H, W = x.shape
    xs = _nhwc(x.float())
    om = torch.empty(H * W, 32, device=dev)
    out = torch.empty(H * W, 32, device=dev)
    src = _src(xs, 32)
    st = _stream(dev)
    capi.check(lib.aarmvs_conv2d(_src_array(src), 1, H, W, 32, 32, 3, 1, _ptr(w[f"pm{level}.w"]), _ptr(w[f"pm{level}.b"]),
                                 _ptr(om), 32, None, st), "aarmvs_conv2d:pm")
    if (math or eng.cell_math) == "tc":
        capi.check(lib.aarmvs_deform_conv_tc(ctypes.byref(src), H, W, _ptr(om), _ptr(w[f"dc{level}.wtc"]), _ptr(w[f"dc{level}.b"]),
                                             _ptr(out), None, st), "aarmvs_deform_conv_tc")
    else:
        capi.check(lib.aarmvs_deform_conv(ctypes.byref(src), H, W, _ptr(om), _ptr(w[f"dc{level}.w"]), _ptr(w[f"dc{level}.b"]),
                                          _ptr(out), None, st), "aarmvs_deform_conv")
    return _nchw(out, H, W)


def loss_grad(model, cost_volume, depth_gt, mask, depth_values, batch=1):
    """d mvsnet_cls_loss / d regularised cost (models/drmvsnet.py:289-291, 347-381) for one sample: the first step of the
    backward pass.  cost_volume [D,H,W] (the train-phase forward's regularised costs, before the softmax), depth_gt / mask
    [H,W], depth_values [D]; batch = B of the loss's mean over samples.  Returns [D,H,W]."""
    _need_cuda(cost_volume)
    eng = model._engine
    D, H, W = cost_volume.shape
    P = H * W
    cost = cost_volume.to(torch.float32).contiguous().view(D, P)
    mk = mask.to(device=cost.device, dtype=torch.float32).reshape(-1).contiguous()
    gt = eng.nearest_hypothesis(depth_values.to(cost.device), depth_gt.to(cost.device), mk.view(H, W)).contiguous()
    scale = 1.0 / ((float(mk.sum().item()) + 1e-6) * batch)
    out = torch.empty_like(cost)
    capi.check(eng.lib.aarmvs_loss_grad(_ptr(cost), _ptr(gt), _ptr(mk), scale, D, P, _ptr(out), _stream(cost.device)), "aarmvs_loss_grad")
    return out.view(D, H, W)

