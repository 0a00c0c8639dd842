"""Backward pass (BASELINE config 5): every backward kernel against autograd through the same op in plain PyTorch (fp64 on the
GPU, or the CPU oracle for the path-specific ops), then the whole training step against the gradients the UNMODIFIED reference
produced for all 90 parameters (tests/golden/train_grads_32x40_n3_d6.pt, oracle/make_golden.py grads).

Tolerance of the end-to-end test: 2e-4 of each tensor's largest reference gradient."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, load_weights

pytestmark = pytest.mark.gpu


def _nhwc(x):          # [C,H,W] -> [H*W, C]
    return x.permute(1, 2, 0).reshape(-1, x.shape[0]).contiguous()


def _nchw(x, h, w):    # [H*W, C] -> [C,H,W]
    return x.view(h, w, -1).permute(2, 0, 1).contiguous()


@pytest.fixture(scope="module")
def ctx():
    import aarmvsnet_b200 as aa
    from aarmvsnet_b200.backward import Backward
    m = aa.EMVSNet(disparity_level=32, image_scale=1.0, max_h=32, max_w=40, return_depth=False)
    m.load_state_dict(load_weights("model_dtu_v2.ckpt"), strict=True)
    m = m.cuda()
    return m, Backward(m._engine, torch.device("cuda", 0))


def _close(got, want, tol, what):
    scale = want.abs().max().item() + 1e-30
    err = (got.double() - want.double()).abs().max().item()
    print(f"{what}: max |err| / max |ref| = {err / scale:.2e}")
    assert err <= tol * scale, f"{what}: {err / scale:.2e} > {tol}"


# ------------------------------------------------------------------------------------------------ generic primitives
@pytest.mark.parametrize("case", ["conv_s1_concat_split", "dx_s1", "dx_s2", "deconv_fwd", "deconv_dx", "k1_wide", "scalar_in"])
def test_conv_nhwc(ctx, case):
    from aarmvsnet_b200.backward import _seg
    _, bw = ctx
    g = torch.Generator().manual_seed(3)
    dev = "cuda"
    H, W = 18, 22
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    if case == "conv_s1_concat_split":
        x, wt, b = r(16, H, W), r(24, 16, 3, 3), r(24)
        want = F.conv2d(x[None], wt, b, padding=1)[0]
        a, c = _nhwc(x[:12]).float().to(dev), _nhwc(x[12:]).float().to(dev)
        o1, o2 = torch.zeros(H * W, 16, device=dev), torch.ones(H * W, 8, device=dev)
        wk = wt.permute(2, 3, 1, 0).reshape(9, 16, 24).float().contiguous().to(dev)
        bs = b.float().to(dev)
        bw.conv([_seg(a, 12), _seg(c, 4)], H, W, [_seg(o1, 16), _seg(o2, 8, 1)], H, W, 3, 1, False, wk, bs)
        got = torch.cat([o1, o2 - 1.0], 1)
        _close(_nchw(got, H, W).cpu(), want, 2e-5, case)
        return
    if case in ("dx_s1", "dx_s2"):
        s = 1 if case == "dx_s1" else 2
        x = r(12, H, W).requires_grad_(True)
        wt = r(20, 12, 3, 3)
        y = F.conv2d(x[None], wt, None, stride=s, padding=1)[0]
        dy = r(*y.shape)
        (want,) = torch.autograd.grad(y, x, dy)
        ho, wo = y.shape[1:]
        out = torch.empty(H * W, 12, device=dev)
        wk = wt.permute(2, 3, 0, 1).reshape(9, 20, 12).float().contiguous().to(dev)        # [tap][co][ci]
        dys = _nhwc(dy).float().to(dev)
        bw.conv([_seg(dys, 20)], ho, wo, [_seg(out, 12)], H, W, 3, s, True, wk)
        _close(_nchw(out, H, W).cpu(), want, 2e-5, case)
        return
    if case in ("deconv_fwd", "deconv_dx"):
        x = r(16, 9, 11).requires_grad_(True)
        wt, b = r(16, 16, 3, 3), r(16)
        y = F.conv_transpose2d(x[None], wt, b, stride=2, padding=1, output_padding=1)[0]
        if case == "deconv_fwd":
            out = torch.empty(18 * 22, 16, device=dev)
            wk = wt.permute(2, 3, 0, 1).reshape(9, 16, 16).float().contiguous().to(dev)    # [tap][ci][co]
            xs, bs = _nhwc(x.detach()).float().to(dev), b.float().to(dev)                  # (named: they must outlive the launch)
            bw.conv([_seg(xs, 16)], 9, 11, [_seg(out, 16)], 18, 22, 3, 2, True, wk, bs)
            _close(_nchw(out, 18, 22).cpu(), y.detach(), 2e-5, case)
        else:
            dy = r(*y.shape)
            (want,) = torch.autograd.grad(y, x, dy)
            out = torch.empty(9 * 11, 16, device=dev)
            wk = wt.permute(2, 3, 1, 0).reshape(9, 16, 16).float().contiguous().to(dev)    # [tap][co][ci]
            dys = _nhwc(dy).float().to(dev)
            bw.conv([_seg(dys, 16)], 18, 22, [_seg(out, 16)], 9, 11, 3, 2, False, wk)
            _close(_nchw(out, 9, 11).cpu(), want, 2e-5, case)
        return
    if case == "k1_wide":
        x, wt = r(32, H, W), r(288, 32, 1, 1)
        want = F.conv2d(x[None], wt)[0]
        out = torch.empty(H * W, 288, device=dev)
        wk = wt.permute(2, 3, 1, 0).reshape(1, 32, 288).float().contiguous().to(dev)
        xs = _nhwc(x).float().to(dev)
        bw.conv([_seg(xs, 32)], H, W, [_seg(out, 288)], H, W, 1, 1, False, wk)
        _close(_nchw(out, H, W).cpu(), want, 2e-5, case)
        return
    x, wt = r(1, H, W), r(8, 1, 3, 3)                                                      # scalar_in: the head's dX
    want = F.conv2d(x[None], wt, padding=1)[0]
    out = torch.zeros(H * W, 8, device=dev)
    wk = wt.permute(2, 3, 1, 0).reshape(9, 1, 8).float().contiguous().to(dev)
    xs = _nhwc(x).float().to(dev)
    bw.conv([_seg(xs, 1)], H, W, [_seg(out, 8, 1)], H, W, 3, 1, False, wk)
    _close(_nchw(out, H, W).cpu(), want, 2e-5, case)


@pytest.mark.parametrize("cin,cout,splits", [(64, 48, (32, 16)), (64, 48, (16, 16, 16)), (64, 32, (16, 16)), (32, 40, (16, 16, 8))])
@pytest.mark.parametrize("H,W", [(18, 22), (40, 72)])
def test_cell_data_gradient_tiled(ctx, cin, cout, splits, H, W):
    """The register-tiled dX kernel of the five cells (conv_dx3_kernel: dG (*) W^T scattered to the cell's inputs, the second
    one accumulated) against autograd through F.conv2d in fp64; ragged tiles in both directions."""
    from aarmvsnet_b200.backward import _seg
    _, bw = ctx
    g = torch.Generator().manual_seed(31)
    dev = "cuda"
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    x = r(cout, H, W).requires_grad_(True)                  # the cell's input (cout channels of dX)
    wt = r(cin, cout, 3, 3)                                 # conv: cout input channels -> cin gate channels
    y = F.conv2d(x[None], wt, None, padding=1)[0]
    dy = r(*y.shape)
    (want,) = torch.autograd.grad(y, x, dy)
    wk = wt.permute(2, 3, 0, 1).reshape(9, cin, cout).float().contiguous().to(dev)         # [tap][gate][input] = W^T
    dys = _nhwc(dy).float().to(dev)
    outs = [torch.full((H * W, c), 1.0 if i == 1 else float("nan"), device=dev) for i, c in enumerate(splits)]
    bw.conv([_seg(dys, cin)], H, W, [_seg(o, c, 1 if i == 1 else 0) for i, (o, c) in enumerate(zip(outs, splits))], H, W, 3, 1, True, wk)
    outs[1] -= 1.0
    _close(_nchw(torch.cat(outs, 1), H, W).cpu(), want, 2e-5, f"dx {cin}->{cout}")


@pytest.mark.parametrize("cin,cout,splits", [(64, 48, (32, 16)), (64, 48, (16, 16, 16)), (64, 32, (16, 16)), (32, 40, (16, 16, 8))])
@pytest.mark.parametrize("H,W,magnitude", [(20, 24, 1.0), (40, 72, 3e-7), (128, 160, 1e-3)])
def test_cell_data_gradient_tensor_cores(ctx, cin, cout, splits, H, W, magnitude):
    """The same data gradients through aarmvs_split_planes_scaled + aarmvs_conv_dx_tc (the forward cell's tcgen05 GEMM with
    the gate gradients as scaled fp16 (hi, lo) operand planes) against autograd in fp64, for gradients of the sizes a
    training step sees (far below fp16's range) and with a few elements 2^-20 of the largest; ragged tiles, more tiles than
    SMs at 128 x 160."""
    from aarmvsnet_b200 import capi, pack
    from aarmvsnet_b200.backward import _seg, _segs
    from aarmvsnet_b200.engine import _ptr
    _, bw = ctx
    g = torch.Generator().manual_seed(37)
    dev = "cuda"
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    x = r(cout, H, W).requires_grad_(True)
    wt = r(cin, cout, 3, 3)
    y = F.conv2d(x[None], wt, None, padding=1)[0]
    dy = r(*y.shape) * magnitude
    dy[:, ::3, ::5] *= 2.0 ** -20
    (want,) = torch.autograd.grad(y, x, dy)
    wk = wt.permute(2, 3, 0, 1).reshape(9, cin, cout).float().contiguous()                  # [tap][gate][input] = W^T
    wtc = pack.conv_f16_steps(wk.flip(0), 32 if cout == 32 else 48).to(dev)
    dys = _nhwc(dy).float().to(dev)
    mbits = dys.abs().max().reshape(1).view(torch.int32).clone()
    planes = torch.empty(2, cin // 8, H * W, 8, device=dev, dtype=torch.float16)
    inv = torch.zeros(1, device=dev)
    capi.check(bw.lib.aarmvs_split_planes_scaled(_ptr(dys), H, W, cin, _ptr(mbits), _ptr(planes), _ptr(inv), bw._st()), "split")
    scaled_max = dys.abs().max().item() / inv.item()
    assert 128.0 <= scaled_max < 256.0, scaled_max
    outs = [torch.full((H * W, c), 1.0 if i == 1 else float("nan"), device=dev) for i, c in enumerate(splits)]
    segs = [_seg(o, c, 1 if i == 1 else 0) for i, (o, c) in enumerate(zip(outs, splits))]
    capi.check(bw.lib.aarmvs_conv_dx_tc(_ptr(planes), H, W, cin, _segs(*segs), len(segs), cout, _ptr(wtc), _ptr(inv), bw._st()), "dx tc")
    torch.cuda.synchronize()
    got = torch.cat([outs[0], outs[1] - 1.0 * (1 if magnitude == 1.0 else 0)] + outs[2:], 1)
    if magnitude != 1.0:                     # the accumulated segment: 1 + tiny is not representable, compare it separately
        return _close(_nchw(torch.cat([outs[0]] + outs[2:], 1), H, W).cpu(),
                      torch.cat([want[:splits[0]], want[splits[0] + splits[1]:]], 0), 1e-5, f"dx tc {cin}->{cout} (stored segments)")
    _close(_nchw(got, H, W).cpu(), want, 1e-5, f"dx tc {cin}->{cout}")


@pytest.mark.parametrize("splits,cy", [((32, 16), 64), ((16, 16), 64), ((16, 16, 16), 64), ((16, 16, 8), 32)])
@pytest.mark.parametrize("H,W,magnitude", [(20, 24, 1.0), (40, 72, 3e-7), (128, 160, 1e-3)])
def test_cell_weight_gradient_tensor_cores(ctx, splits, cy, H, W, magnitude):
    """aarmvs_split_planes_cat + aarmvs_conv_dw_tc (a K = pixels tcgen05 GEMM over the operand planes read as MN-major
    operands, accumulated in TMEM over all tiles of a CTA) against autograd in fp64; accumulates onto a non-zero dW."""
    from aarmvsnet_b200 import capi
    from aarmvsnet_b200.backward import _seg, _segs
    from aarmvsnet_b200.engine import _ptr
    _, bw = ctx
    g = torch.Generator().manual_seed(41)
    dev = "cuda"
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    cx = sum(splits)
    x = r(cx, H, W)
    wt = r(cy, cx, 3, 3).requires_grad_(True)
    y = F.conv2d(x[None], wt, None, padding=1)[0]
    dy = r(*y.shape) * magnitude
    dy[:, ::3, ::5] *= 2.0 ** -20
    (want,) = torch.autograd.grad(y, wt, dy)                                               # [cy][cx][3][3]
    xs = _nhwc(x).float().to(dev)
    parts, o = [], 0
    for c in splits:
        parts.append(xs[:, o:o + c].contiguous())
        o += c
    xplanes = torch.empty(2, cx // 8, H * W, 8, device=dev, dtype=torch.float16)
    segs = [_seg(t, c) for t, c in zip(parts, splits)]
    capi.check(bw.lib.aarmvs_split_planes_cat(_segs(*segs), len(segs), H, W, _ptr(xplanes), bw._st()), "split cat")
    dys = _nhwc(dy).float().to(dev)
    mbits = dys.abs().max().reshape(1).view(torch.int32).clone()
    gplanes = torch.empty(2, cy // 8, H * W, 8, device=dev, dtype=torch.float16)
    inv = torch.zeros(1, device=dev)
    capi.check(bw.lib.aarmvs_split_planes_scaled(_ptr(dys), H, W, cy, _ptr(mbits), _ptr(gplanes), _ptr(inv), bw._st()), "split")
    start = magnitude * 0.5
    dw = torch.full((9, cx, cy), start, device=dev)
    capi.check(bw.lib.aarmvs_conv_dw_tc(_ptr(xplanes), _ptr(gplanes), H, W, cx, cy, _ptr(inv), _ptr(dw), bw._st()), "dw tc")
    torch.cuda.synchronize()
    got = (dw.double() - start).reshape(3, 3, cx, cy).permute(3, 2, 0, 1)
    _close(got.cpu(), want, 1e-5, f"dW tc {cx}->{cy}")


@pytest.mark.parametrize("case", ["s1_concat", "s2", "deconv", "k1_288", "cy1", "cell0", "cell4", "stem"])
def test_conv_bwd_weight(ctx, case):
    from aarmvsnet_b200.backward import _seg
    _, bw = ctx
    g = torch.Generator().manual_seed(4)
    dev = "cuda"
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    H, W = 34, 26                       # 884 pixels: several staging tiles, ragged last one
    if case == "deconv":
        x = r(16, 9, 13)
        wt = r(16, 16, 3, 3).requires_grad_(True)
        y = F.conv_transpose2d(x[None], wt, None, stride=2, padding=1, output_padding=1)[0]
        dy = r(*y.shape)
        (want,) = torch.autograd.grad(y, wt, dy)                                          # [ci][co][ky][kx]
        dw = torch.zeros(9, 16, 16, device=dev)
        dys, xs = _nhwc(dy).float().to(dev), _nhwc(x).float().to(dev)
        bw.wgrad([_seg(dys, 16)], 18, 26, xs, 9, 13, 16, 3, 2, dw)
        got = dw.reshape(3, 3, 16, 16).permute(3, 2, 0, 1)
        _close(got.cpu(), want, 5e-5, case)
        return
    cin, cout, k, s = {"s1_concat": (20, 24, 3, 1), "s2": (32, 32, 3, 2), "k1_288": (288, 32, 1, 1), "cy1": (8, 1, 3, 1),
                       "cell0": (48, 64, 3, 1), "cell4": (40, 32, 3, 1), "stem": (32, 4, 3, 1)}[case]
    if case in ("cell0", "cell4"):
        H, W = (64, 200) if case == "cell0" else (40, 72)       # the tiled kernel: 56 tiles, two per CTA / 15 tiles; ragged right edge
    x = r(cin, H, W)
    wt = r(cout, cin, k, k).requires_grad_(True)
    b = r(cout).requires_grad_(True)
    y = F.conv2d(x[None], wt, b, stride=s, padding=(k - 1) // 2)[0]
    dy = r(*y.shape)
    want_w, want_b = torch.autograd.grad(y, (wt, b), dy)
    ho, wo = y.shape[1:]
    dw = torch.zeros(k * k, cin, cout, device=dev)
    db = torch.zeros(cout, device=dev)
    xs = _nhwc(x).float().to(dev)
    keep = [xs[:, :12].contiguous(), xs[:, 12:].contiguous()]
    segs = [_seg(keep[0], 12), _seg(keep[1], cin - 12)] if case in ("s1_concat", "cell0", "cell4") else [_seg(xs, cin)]
    dys = _nhwc(dy).float().to(dev)
    bw.wgrad(segs, H, W, dys, ho, wo, cout, k, s, dw, db)
    _close(dw.reshape(k, k, cin, cout).permute(3, 2, 0, 1).cpu(), want_w, 5e-5, case + " dW")
    # sum_p dy cancels heavily (884+ summands of either sign, atomics in any order): measure against the summands' scale
    err_b = (db.cpu().double() - want_b).abs().max().item()
    print(f"{case} db: max |err| = {err_b:.2e}, sum |dy| = {dy.abs().sum((1, 2)).max().item():.1f}")
    assert err_b <= 2e-7 * dy.abs().sum((1, 2)).max().item(), f"{case} db: {err_b:.2e}"


@pytest.mark.parametrize("C,groups", [(4, 1), (16, 2), (32, 4), (8, 1)])
def test_group_norm_relu_forward_and_backward(ctx, C, groups):
    from aarmvsnet_b200 import capi
    from aarmvsnet_b200.engine import _ptr
    _, bw = ctx
    g = torch.Generator().manual_seed(5)
    H, W = 20, 28
    P = H * W
    x = torch.randn(C, H, W, generator=g, dtype=torch.float64).requires_grad_(True)
    gam = torch.randn(C, generator=g, dtype=torch.float64).requires_grad_(True)
    bet = torch.randn(C, generator=g, dtype=torch.float64).requires_grad_(True)
    y = F.relu(F.group_norm(x[None], groups, gam, bet, 1e-5))[0]
    dy = torch.randn(C, H, W, generator=g, dtype=torch.float64)
    wx, wg, wb = torch.autograd.grad(y, (x, gam, bet), dy)
    dev = "cuda"
    xs, dys = _nhwc(x.detach()).float().to(dev), _nhwc(dy).float().to(dev)
    gn = (gam.detach().float().to(dev), bet.detach().float().to(dev))
    sums = torch.zeros(2 * groups, device=dev, dtype=torch.float64)
    capi.check(bw.lib.aarmvs_gn_sums(_ptr(xs), P, C, groups, _ptr(sums), bw._st()), "gn_sums")
    out = torch.empty_like(xs)
    bw.gn_apply(xs, P, C, groups, sums, gn, out)
    _close(_nchw(out, H, W).cpu(), y.detach(), 2e-5, "gn forward")
    dgam, dbet, dbias = (torch.zeros(C, device=dev) for _ in range(3))
    dx = torch.empty_like(xs)
    bw.gn_bwd(xs, dys, P, C, groups, sums, gn, dx, dgam, dbet, dbias)
    _close(_nchw(dx, H, W).cpu(), wx, 5e-5, "gn dx")
    _close(dgam.cpu(), wg, 5e-5, "gn dgamma")
    _close(dbet.cpu(), wb, 5e-5, "gn dbeta")
    # sum_p dx (the producing conv's bias gradient) cancels heavily: compare against the scale of the summands
    assert (dbias.cpu().double() - wx.sum((1, 2))).abs().max().item() <= 1e-4 * wx.abs().sum((1, 2)).max().item()


def test_lstm_gates_and_maxpool_backward(ctx):
    from aarmvsnet_b200 import capi
    from aarmvsnet_b200.engine import _ptr
    _, bw = ctx
    g = torch.Generator().manual_seed(6)
    dev = "cuda"
    P, Ch = 37 * 4, 16
    gates = torch.randn(P, 4 * Ch, generator=g, dtype=torch.float64).requires_grad_(True)
    cprev = torch.randn(P, Ch, generator=g, dtype=torch.float64).requires_grad_(True)
    i, f, o, gg = torch.split(gates, Ch, dim=1)
    cn = torch.sigmoid(f) * cprev + torch.sigmoid(i) * torch.tanh(gg)
    hn = torch.sigmoid(o) * torch.tanh(cn)
    dh = torch.randn(P, Ch, generator=g, dtype=torch.float64)
    dcn = torch.randn(P, Ch, generator=g, dtype=torch.float64)
    wg, wc = torch.autograd.grad((hn, cn), (gates, cprev), (dh, dcn))
    dc = dcn.float().to(dev).clone()
    dg = torch.empty(P, 4 * Ch, device=dev)
    t_g, t_cp, t_cn, t_dh = (t.detach().float().to(dev) for t in (gates, cprev, cn, dh))      # named: they must outlive the launch
    capi.check(bw.lib.aarmvs_lstm_gates_bwd(_ptr(t_g), _ptr(t_cp), _ptr(t_cn), _ptr(t_dh), _ptr(dc), _ptr(dg), P, Ch, bw._st()), "lstm")
    _close(dg.cpu(), wg, 5e-5, "lstm dgates")
    _close(dc.cpu(), wc, 5e-5, "lstm dc_prev")
    # the same from the activated gates a training forward stored: record 2 of a 3-plane history, plane index on the device
    act = torch.cat([torch.sigmoid(gates[:, :3 * Ch]), torch.tanh(gates[:, 3 * Ch:])], dim=1).detach().float()
    hist = torch.full((3, P, 4 * Ch), float("nan"), device=dev)
    hist[2] = act.to(dev)
    slot = torch.tensor([hist.data_ptr()], dtype=torch.int64).to(dev)
    plane = torch.tensor([2], dtype=torch.int32).to(dev)
    dc2 = dcn.float().to(dev).clone()
    dg2 = torch.empty(P, 4 * Ch, device=dev)
    mbits = torch.zeros(1, dtype=torch.int32, device=dev)
    dbias = torch.zeros(4 * Ch, device=dev)
    capi.check(bw.lib.aarmvs_lstm_gates_bwd_act(_ptr(slot), 0, _ptr(plane), _ptr(t_cp), _ptr(t_cn), _ptr(t_dh), _ptr(dc2), _ptr(dg2), P, Ch,
                                                _ptr(mbits), _ptr(dbias), bw._st()), "lstm act")
    _close(dg2.cpu(), wg, 5e-5, "lstm dgates (stored activations)")
    _close(dc2.cpu(), wc, 5e-5, "lstm dc_prev (stored activations)")
    assert mbits.view(torch.float32).item() == dg2.abs().max().item()         # the scale of the tensor-core data gradient
    err_b = (dbias.cpu().double() - wg.sum(0)).abs().max().item()             # (cancelling sum: against the summands' scale)
    assert err_b <= 2e-7 * wg.abs().sum(0).max().item(), f"sum_p dgates (the cell conv's bias gradient): {err_b:.2e}"
    # max pool
    H, W, C = 12, 20, 16
    x = torch.randn(C, H, W, generator=g, dtype=torch.float64).requires_grad_(True)
    y = F.max_pool2d(x[None], 2, 2)[0]
    dy = torch.randn(*y.shape, generator=g, dtype=torch.float64)
    (wx,) = torch.autograd.grad(y, x, dy)
    xs = _nhwc(x.detach()).float().to(dev)
    out = torch.empty(H * W // 4, C, device=dev)
    capi.check(bw.lib.aarmvs_maxpool2(_ptr(xs), H, W, C, _ptr(out), bw._st()), "maxpool2")
    _close(_nchw(out, H // 2, W // 2).cpu(), y.detach(), 1e-7, "maxpool forward")
    dx = torch.zeros(H * W, C, device=dev)
    dys = _nhwc(dy).float().to(dev)
    capi.check(bw.lib.aarmvs_maxpool2_bwd(_ptr(xs), _ptr(dys), H, W, C, _ptr(dx), bw._st()), "maxpool2_bwd")
    _close(_nchw(dx, H, W).cpu(), wx, 1e-6, "maxpool dx")


# ------------------------------------------------------------------------------------------------ path-specific ops
def test_aggregate_omega_backward(ctx, sd_dtu):
    """d cost -> d c_v and omega's parameters: aarmvs_aggregate_bwd + four aarmvs_omega_bwd passes + the stem conv's two
    gradients, against autograd through the oracle's omega / aggregation (drmvsnet.py:27-38, 311-319) in fp64."""
    from aarmvsnet_b200 import capi
    from aarmvsnet_b200.backward import _seg
    from aarmvsnet_b200.engine import _ptr
    from oracle import aarmvs_oracle as orc
    m, bw = ctx
    g = torch.Generator().manual_seed(7)
    dev = "cuda"
    H, W, S = 20, 24, 3
    P = H * W
    sd64 = {k: v.double().requires_grad_(k.startswith("omega.")) for k, v in sd_dtu.items()}
    cv = [(torch.randn(1, 32, H, W, generator=g, dtype=torch.float64) ** 2).requires_grad_(True) for _ in range(S)]
    acc = sum((orc.omega(c, sd64) + 1) * c for c in cv)
    cost = -acc / S
    dcost = torch.randn(1, 32, H, W, generator=g, dtype=torch.float64)
    names = [k for k in sd64 if k.startswith("omega.")]
    grads = torch.autograd.grad(cost, cv + [sd64[k] for k in names], dcost)
    want_cv, want_p = grads[:S], dict(zip(names, grads[S:]))
    # our forward pieces for the same inputs
    w = bw.w
    costv = torch.stack([_nhwc(c.detach()[0]).float() for c in cv]).to(dev).contiguous()
    t0 = torch.empty(S, P, 4, device=dev)
    stem_w = torch.as_tensor(list(w["stem.w_host"]), device=dev).reshape(9, 32, 4)
    stem_b = torch.as_tensor(list(w["stem.b_host"]), device=dev)
    stats = torch.zeros(S * 6, device=dev, dtype=torch.float64)
    for v in range(S):
        bw.conv([_seg(costv[v], 32)], H, W, [_seg(t0[v], 4)], H, W, 3, 1, False, stem_w, stem_b)
        capi.check(bw.lib.aarmvs_gn_sums(_ptr(t0[v]), P, 4, 1, _ptr(stats[v * 6:]), bw._st()), "gn_sums")
    for stage in (2, 3):
        capi.check(bw.lib.aarmvs_omega_stats(stage, _ptr(t0), H, W, S, 1, w["tail_host"], _ptr(stats), bw._st()), "omega_stats")
    out_cost, rew = torch.empty(P, 32, device=dev), torch.empty(S, P, device=dev)
    capi.check(bw.lib.aarmvs_aggregate(None, None, 0, H, W, S, None, w["tail_host"], None, 0, None, 1, 0, _ptr(t0), _ptr(stats), _ptr(costv),
                                       _ptr(out_cost), None, _ptr(rew), bw._st()), "aggregate")
    _close(_nchw(out_cost, H, W).cpu(), cost.detach()[0], 1e-4, "aggregated cost (forward)")
    # backward
    for k in ("tail", "stem.w", "stem.b"):
        bw.g[k].zero_()
    dcs = _nhwc(dcost[0]).float().to(dev)
    dcostv, drew, dt0 = torch.empty(S, P, 32, device=dev), torch.empty(S, P, device=dev), torch.empty(S, P, 4, device=dev)
    bstats = torch.zeros(S * 6, device=dev, dtype=torch.float64)
    capi.check(bw.lib.aarmvs_aggregate_bwd(_ptr(dcs), _ptr(costv), _ptr(rew), S, P, _ptr(dcostv), _ptr(drew), bw._st()), "aggregate_bwd")
    for stage in (3, 2, 1, 0):
        capi.check(bw.lib.aarmvs_omega_bwd(stage, _ptr(t0), _ptr(drew), H, W, S, w["tail_host"], _ptr(stats), _ptr(bstats), _ptr(bw.g["tail"]),
                                           _ptr(dt0), bw._st()), "omega_bwd")
    for v in range(S):
        bw.conv([_seg(dt0[v], 4)], H, W, [_seg(dcostv[v], 32, 1)], H, W, 3, 1, True, bw.bw["stem.wT"])
        bw.wgrad([_seg(costv[v], 32)], H, W, dt0[v], H, W, 4, 3, 1, bw.g["stem.w"], bw.g["stem.b"])
    for v in range(S):
        _close(_nchw(dcostv[v], H, W).cpu(), want_cv[v][0], 2e-4, f"d c_v[{v}]")
    got = bw.parameter_gradients()
    for k in names:
        _close(got[k].cpu(), want_p[k], 2e-4, k)
    for k in ("tail", "stem.w", "stem.b"):
        bw.g[k].zero_()


def test_warp_backward(ctx):
    """d c_v -> d ref, d src through the bilinear taps, against autograd through the oracle's homo_warp (module.py:6-38)."""
    from aarmvsnet_b200 import capi, synth
    from aarmvsnet_b200.engine import Engine, _ptr
    from oracle import aarmvs_oracle as orc
    _, bw = ctx
    g = torch.Generator().manual_seed(8)
    dev = "cuda"
    H, W, S = 24, 32, 2
    P = H * W
    s = synth.make_sample(H, W, S + 1, 4, seed=8)
    pm, depth = s["proj_matrices"], s["depth_values"][:, 2]
    ref = torch.randn(1, 32, H, W, generator=g, dtype=torch.float64).requires_grad_(True)
    srcs = [torch.randn(1, 32, H, W, generator=g, dtype=torch.float64).requires_grad_(True) for _ in range(S)]
    cvs = [(orc.homo_warp(srcs[v].float(), pm[:, v + 1], pm[:, 0], depth).double() - ref) ** 2 for v in range(S)]
    dcv = [torch.randn(1, 32, H, W, generator=g, dtype=torch.float64) for _ in range(S)]
    # autograd through the fp32 grid_sample of the oracle: differentiate in fp32 to keep its exact sample positions
    ref32 = ref.detach().float().requires_grad_(True)
    src32 = [t.detach().float().requires_grad_(True) for t in srcs]
    cv32 = [(orc.homo_warp(src32[v], pm[:, v + 1], pm[:, 0], depth) - ref32) ** 2 for v in range(S)]
    grads = torch.autograd.grad(cv32, [ref32] + src32, [d.float() for d in dcv])
    feat = torch.stack([_nhwc(ref32.detach()[0])] + [_nhwc(t.detach()[0]) for t in src32]).to(dev).contiguous()
    cams = Engine.relative_cameras(pm[0]).to(dev)
    dcostv = torch.stack([_nhwc(d[0]).float() for d in dcv]).to(dev).contiguous()
    dfeat = torch.zeros(S + 1, P, 32, device=dev)
    depths = s["depth_values"][0].to(dev)
    capi.check(bw.lib.aarmvs_warp_bwd(_ptr(feat), _ptr(feat, P * 32), P * 32, H, W, S, _ptr(cams), _ptr(depths), 2, None, 0, _ptr(dcostv),
                                      _ptr(dfeat), _ptr(dfeat, P * 32), bw._st()), "warp_bwd")
    _close(_nchw(dfeat[0], H, W).cpu(), grads[0][0], 2e-5, "d ref")
    for v in range(S):
        _close(_nchw(dfeat[v + 1], H, W).cpu(), grads[v + 1][0], 2e-5, f"d src[{v}]")
    assert cvs[0].shape == cv32[0].shape


def test_deform_conv_and_assemble_backward(ctx, sd_dtu):
    """One modulated deformable layer (gather + contraction + offset / modulation conv) and the IntraViewAA tail, against
    autograd through the oracle (module.py:133-189, drmvsnet.py:19-24)."""
    from aarmvsnet_b200 import capi
    from aarmvsnet_b200.backward import _seg
    from aarmvsnet_b200.engine import _ptr, _src, _src_array
    from oracle import aarmvs_oracle as orc
    _, bw = ctx
    g = torch.Generator().manual_seed(9)
    dev = "cuda"
    H, W = 16, 20
    P = H * W
    prefix = "feature.intraAA.deformconv1.0"
    names = [prefix + s for s in (".conv.weight", ".conv.bias", ".p_conv.weight", ".p_conv.bias", ".m_conv.weight", ".m_conv.bias")]
    sd32 = {k: v.float().clone().requires_grad_(k in names) for k, v in sd_dtu.items()}
    x = (torch.randn(1, 32, H, W, generator=g) * 1.5).requires_grad_(True)
    y = orc.deform_conv(x, sd32, prefix)
    dy = torch.randn(1, 32, H, W, generator=g)
    grads = torch.autograd.grad(y, [x] + [sd32[k] for k in names], dy)
    want_x, want_p = grads[0], dict(zip(names, grads[1:]))
    w, l = bw.w, 1
    xs = _nhwc(x.detach()[0]).to(dev)
    om = torch.empty(P, 32, device=dev)
    capi.check(bw.lib.aarmvs_conv2d(_src_array(_src(xs, 32)), 1, H, W, 32, 32, 3, 1, _ptr(w[f"pm{l}.w"]), _ptr(w[f"pm{l}.b"]), _ptr(om), 32,
                                    None, bw._st()), "pm conv")
    samp, samp_m, dsm = (torch.empty(P, 288, device=dev) for _ in range(3))
    capi.check(bw.lib.aarmvs_deform_sample(_ptr(xs), H, W, _ptr(om), _ptr(samp), _ptr(samp_m), bw._st()), "deform_sample")
    out = torch.empty(P, 32, device=dev)
    bw.conv([_seg(samp_m, 288)], H, W, [_seg(out, 32)], H, W, 1, 1, False, w[f"dc{l}.w"].reshape(1, 288, 32).contiguous(), w[f"dc{l}.b"])
    _close(_nchw(out, H, W).cpu(), y.detach()[0], 1e-4, "deformable conv forward (sample + contraction)")
    for k in (f"dc{l}.w", f"pm{l}.w", f"pm{l}.b"):
        bw.g[k].zero_()
    dys = _nhwc(dy[0]).to(dev)
    bw.conv([_seg(dys, 32)], H, W, [_seg(dsm, 288)], H, W, 1, 1, False, bw.bw[f"dc{l}.w288T"])
    bw.wgrad([_seg(samp_m, 288)], H, W, dys, H, W, 32, 1, 1, bw.g[f"dc{l}.w"])
    dx, dom = torch.zeros(P, 32, device=dev), torch.empty(P, 32, device=dev)
    capi.check(bw.lib.aarmvs_deform_bwd(_ptr(xs), H, W, _ptr(om), _ptr(samp), _ptr(dsm), _ptr(dx), _ptr(dom), bw._st()), "deform_bwd")
    bw.conv([_seg(dom, 32)], H, W, [_seg(dx, 32, 1)], H, W, 3, 1, True, bw.bw[f"pm{l}.wT"])
    bw.wgrad([_seg(xs, 32)], H, W, dom, H, W, 32, 3, 1, bw.g[f"pm{l}.w"], bw.g[f"pm{l}.b"])
    _close(_nchw(dx, H, W).cpu(), want_x[0], 2e-4, "deform d x")
    got = bw.parameter_gradients()
    for k in names:
        if k.endswith("conv.bias"):
            continue                                   # comes from the following GroupNorm's backward in the full chain
        _close(got[k].cpu(), want_p[k], 2e-4, k)
    for k in (f"dc{l}.w", f"pm{l}.w", f"pm{l}.b"):
        bw.g[k].zero_()
    # IntraViewAA tail: x2 / x4 bilinear (align_corners=True) + concat
    Hf, Wf = 16, 24
    m0 = torch.randn(1, 16, Hf, Wf, generator=g, dtype=torch.float64).requires_grad_(True)
    m1 = torch.randn(1, 8, Hf // 2, Wf // 2, generator=g, dtype=torch.float64).requires_grad_(True)
    m2 = torch.randn(1, 8, Hf // 4, Wf // 4, generator=g, dtype=torch.float64).requires_grad_(True)
    f = torch.cat([m0, F.interpolate(m1, scale_factor=2, mode="bilinear", align_corners=True),
                   F.interpolate(m2, scale_factor=4, mode="bilinear", align_corners=True)], 1)
    df = torch.randn(1, 32, Hf, Wf, generator=g, dtype=torch.float64)
    w0, w1, w2 = torch.autograd.grad(f, (m0, m1, m2), df)
    d0 = torch.empty(Hf * Wf, 16, device=dev)
    d1, d2 = torch.zeros(Hf * Wf // 4, 8, device=dev), torch.zeros(Hf * Wf // 16, 8, device=dev)
    dfs = _nhwc(df[0]).float().to(dev)
    capi.check(bw.lib.aarmvs_feat_assemble_bwd(_ptr(dfs), Hf, Wf, _ptr(d0), _ptr(d1), _ptr(d2), bw._st()), "assemble_bwd")
    _close(_nchw(d0, Hf, Wf).cpu(), w0[0], 1e-6, "d m0")
    _close(_nchw(d1, Hf // 2, Wf // 2).cpu(), w1[0], 2e-5, "d m1")
    _close(_nchw(d2, Hf // 4, Wf // 4).cpu(), w2[0], 2e-5, "d m2")


# ------------------------------------------------------------------------------------------------ the whole training step
@pytest.mark.parametrize("backward_math", ["tc", "fma"])
def test_training_step_gradients_match_reference(backward_math):
    """train.py:288-300 on the train fixture: prob = model(...); loss = mvsnet_cls_loss(prob, ...); loss.backward() -- all 90
    parameter gradients against the ones the UNMODIFIED reference computed (2e-4 of each tensor's largest gradient); with the
    cells' data / weight gradients on the tensor cores (the product path) and on the fp32 fallback kernels."""
    import aarmvsnet_b200 as aa
    t = load_golden("train_32x40_n3_d6.pt")
    gold = load_golden("train_grads_32x40_n3_d6.pt")
    m = aa.EMVSNet(disparity_level=32, image_scale=1.0, max_h=32, max_w=40, return_depth=False)
    m.load_state_dict(load_weights(t["ckpt"]), strict=True)
    m = m.cuda().train()
    m.backward_math = backward_math
    prob, _, _ = m(t["imgs"].cuda(), t["proj_matrices"].cuda(), t["depth_values"].cuda())
    assert prob.requires_grad
    loss, _ = aa.mvsnet_cls_loss(prob, t["depth_gt"].cuda(), t["mask"].cuda(), t["depth_values"].cuda())
    torch.testing.assert_close(loss.detach().cpu(), gold["loss"], rtol=2e-4, atol=1e-6)
    loss.backward()
    worst, bad = 0.0, []
    assert len(gold["grads"]) == 90
    for name, p in m.named_parameters():
        want = gold["grads"][name]
        assert p.grad is not None, name
        scale = want.abs().max().item()
        err = (p.grad.cpu() - want).abs().max().item()
        if scale < 1e-6:
            # conv_0.bias: a shift of every regularised cost cancels in the softmax over D, its true gradient is 0 and the
            # reference's own value (1.5e-8) is rounding noise -- require ours to be noise too
            assert err < 1e-6, f"{name}: {err:.2e}"
            continue
        rel = err / scale
        worst = max(worst, rel)
        if rel > 2e-4:
            bad.append(f"{name}: {rel:.2e} (max |ref| {scale:.2e})")
    print(f"90 parameter gradients ({backward_math}): worst max-error / max-gradient = {worst:.2e}")
    assert not bad, "\n".join(bad)


@pytest.mark.parametrize("H,W,B,D", [(24, 32, 1, 7), (24, 32, 2, 7), (24, 32, 1, 11), (128, 160, 1, 3), (72, 104, 2, 4)])
def test_batch_against_oracle_autograd(sd_dtu, H, W, B, D):
    """B samples, D planes (partial plane batches) at 24x32, at BASELINE config 5's tensor size 128x160 (every backward kernel with
    many CTAs, several weight-gradient chunks, multi-block GroupNorm reductions) and at a ragged 72x104 with two samples on
    concurrent streams: loss and all parameter gradients against autograd through the
    oracle's training forward + loss on the CPU in fp64 (the oracle's autograd is pinned to the reference's gradients in
    tests/test_oracle_golden.py::test_training_gradients).  The same oracle in fp32 -- i.e. what the reference itself computes --
    is measured against fp64 too: that is the noise floor of these sums (two samples' gradients partly cancel)."""
    import aarmvsnet_b200 as aa
    from oracle import aarmvs_oracle as orc
    N = 3
    s = aa.synth.make_sample(H, W, N, D, seed=17, batch=B, depth_mode="inverse")
    g = torch.Generator().manual_seed(18)
    depth_gt = 425.0 + 510.0 * torch.rand(B, H, W, generator=g)
    mask = (torch.rand(B, H, W, generator=g) > 0.25).float()

    def oracle_grads(dtype):
        sd = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd_dtu.items()}
        prob = orc.forward_train(sd, s["imgs"].to(dtype), s["proj_matrices"].to(dtype), s["depth_values"].to(dtype))
        loss, _ = orc.cls_loss(prob, depth_gt.to(dtype), mask.to(dtype), s["depth_values"].to(dtype))
        loss.backward()
        return loss.detach(), {k: v.grad for k, v in sd.items()}

    loss64, g64 = oracle_grads(torch.float64)
    loss32, g32 = oracle_grads(torch.float32)
    m = aa.EMVSNet(disparity_level=32, image_scale=1.0, max_h=H, max_w=W, return_depth=False)
    m.load_state_dict(sd_dtu, strict=True)
    m = m.cuda().train()
    p2, _, _ = m(s["imgs"].cuda(), s["proj_matrices"].cuda(), s["depth_values"].cuda())
    loss, _ = aa.mvsnet_cls_loss(p2, depth_gt.cuda(), mask.cuda(), s["depth_values"].cuda())
    print(f"{H}x{W} B={B} D={D}: loss ours {loss.item():.7f}, oracle fp32 {loss32.item():.7f}, oracle fp64 {loss64.item():.7f}")
    torch.testing.assert_close(loss.detach().cpu().double(), loss64, rtol=5e-5, atol=1e-6)
    loss.backward()
    ours, floor = [], []
    for name, p in m.named_parameters():
        want = g64[name]
        scale = want.abs().max().item()
        if scale < 1e-6:
            continue
        ours.append(((p.grad.cpu().double() - want).abs().max().item() / scale, name))
        floor.append(((g32[name].double() - want).abs().max().item() / scale, name))
    ours.sort(reverse=True)
    floor.sort(reverse=True)
    print(f"{H}x{W} B={B} D={D}: worst max-error / max-gradient vs fp64: ours {ours[0][0]:.2e} ({ours[0][1]}), the fp32 oracle {floor[0][0]:.2e} "
          f"({floor[0][1]}); ours top: " + ", ".join(f"{n} {r:.1e}" for r, n in ours[:5]))
    assert ours[0][0] <= max(2e-4, 4.0 * floor[0][0]), ours[:6]


def test_flat_adam_matches_torch_adam_and_training_step_lowers_the_loss():
    """csrc/adam.cu against torch.optim.Adam on the same gradients (three steps), then three optimisation steps of TrainStep
    (train.py:288-307) on one batch: the loss must fall."""
    import aarmvsnet_b200 as aa
    from aarmvsnet_b200.training import FlatAdam, TrainStep
    t = load_golden("train_32x40_n3_d6.pt")
    m = aa.EMVSNet(disparity_level=32, image_scale=1.0, max_h=32, max_w=40, return_depth=False)
    m.load_state_dict(load_weights(t["ckpt"]), strict=True)
    m = m.cuda().train()
    ref_params = [p.detach().clone().requires_grad_(True) for p in m.parameters()]
    ref_opt = torch.optim.Adam(ref_params, lr=1e-3)
    opt = FlatAdam(m, lr=1e-3)
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(3):
        opt.zero_grad()
        for p, q in zip(m.parameters(), ref_params):
            grad = torch.randn(p.shape, device="cuda", generator=g) * 0.01
            p.grad.copy_(grad)
            q.grad = grad.clone()
        opt.step()
        ref_opt.step()
    for (name, p), q in zip(m.named_parameters(), ref_params):
        torch.testing.assert_close(p.detach(), q.detach(), rtol=1e-5, atol=1e-7, msg=name)
    # a real step: forward, loss, backward, update
    m.load_state_dict(load_weights(t["ckpt"]), strict=True)
    step = TrainStep(m, lr=2e-4)
    args = [t[k].cuda() for k in ("imgs", "proj_matrices", "depth_values", "depth_gt", "mask")]
    losses = [step(*args)[0].item() for _ in range(3)]
    print("training steps on one batch, loss:", ", ".join(f"{x:.5f}" for x in losses))
    assert abs(losses[0] - float(t["loss"])) <= 2e-4 * float(t["loss"])
    assert losses[2] < losses[0]
