"""Parity beyond the golden fixtures: a mid-size case against the CPU oracle run on the spot, and size-independent
properties at BASELINE.json's full configuration (cfg2: 1600x1184, N=7, D=512) where the oracle would need hours."""
import pytest
import torch

from conftest import load_weights

pytestmark = pytest.mark.gpu

# plane-mismatch bar against the oracle: 4x the reference's own fp32-vs-fp64 disagreement (4.9e-5); measured
# values in round 1 were 2.4e-5 (cfg1 crop) and 5.3e-6 (cfg2 resolution)
MISMATCH_BAR = 2e-4


def _model(h, w, ckpt="model_dtu_v2.ckpt", **kw):
    import aarmvsnet_b200 as aa
    m = aa.EMVSNet(disparity_level=32, image_scale=1.0, max_h=h, max_w=w, return_depth=True, **kw)
    m.load_state_dict(load_weights(ckpt), strict=True)
    return m.cuda().eval()


def _run(m, s):
    with torch.no_grad():
        return m(s["imgs"].cuda(), s["proj_matrices"].cuda(), s["depth_values"].cuda())


@pytest.mark.parametrize("smooth", [False, True])
def test_mid_size_against_oracle(smooth):
    """104x136 (ragged against the 16x8 / 8x16 / 8x32 kernel tiles), N=5, D=20, inverse depths."""
    from aarmvsnet_b200 import synth
    from oracle import aarmvs_oracle as orc
    s = synth.make_sample(104, 136, 5, 20, seed=21, depth_mode="inverse", smooth=smooth)
    out = _run(_model(104, 136), s)
    ref = orc.forward_test(load_weights("model_dtu_v2.ckpt"), s["imgs"], s["proj_matrices"], s["depth_values"], return_index=True)
    same = out["index"].cpu().long() == ref["index"]
    frac = same.float().mean().item()
    print(f"smooth={smooth}: same plane on {frac * 100:.4f}% of pixels")
    assert frac >= 1.0 - MISMATCH_BAR
    torch.testing.assert_close(out["depth"].cpu()[same], ref["depth"][same], rtol=0, atol=0)
    torch.testing.assert_close(out["photometric_confidence"].cpu()[same], ref["photometric_confidence"][same], rtol=1e-3, atol=1e-3)


def _compare(tag, out, ref, s, conf_tol=1e-3):
    """north_star parity statistics of one forward against the oracle: plane-mismatch fraction (bar: MISMATCH_BAR, next to the
    reference's own fp32-vs-fp64 floor of 4.9e-5), neighbouring planes only, depth identical where the plane is, depth error
    against the range, confidence."""
    idx = out["index"].cpu().long()
    same = idx == ref["index"]
    mismatch = 1.0 - same.float().mean().item()
    dist = (idx - ref["index"]).abs().max().item()
    rng = (s["depth_values"].max() - s["depth_values"].min()).item()
    err = (out["depth"].cpu() - ref["depth"]).abs()
    dconf = (out["photometric_confidence"].cpu() - ref["photometric_confidence"]).abs()
    print(f"{tag}: plane mismatch fraction {mismatch:.2e} (reference fp32 vs fp64: 4.9e-05) on {same.numel()} pixels, max index "
          f"distance {dist}, max |depth error| / range {err.max().item() / rng:.2e}, max |dconf| {dconf[same].max().item():.2e}")
    assert mismatch <= MISMATCH_BAR
    # the differing pixels are near ties of the ORACLE's own regularised cost (with white-noise images the cost over depth is
    # multi-modal, so the two tied planes need not be neighbours: cfg1 has 2 such pixels of 327 680, 15 planes apart)
    if "cost_reg" in ref and (~same).any():
        c = ref["cost_reg"]                                           # [B,D,H,W]
        gap = (c.gather(1, ref["index"].unsqueeze(1)) - c.gather(1, idx.unsqueeze(1))).squeeze(1)[~same]
        print(f"{tag}: oracle cost gap between its plane and ours at the {int((~same).sum())} differing pixels: max {gap.abs().max().item():.2e}")
        assert gap.abs().max().item() <= 2e-3
    else:
        assert dist <= 1 or mismatch <= 2e-5
    assert (err[same] == 0).all()
    interval = (s["depth_values"][0, 1:] - s["depth_values"][0, :-1]).abs()
    assert (err <= 0.5 * interval.min()).float().mean().item() >= 1.0 - MISMATCH_BAR
    torch.testing.assert_close(out["photometric_confidence"].cpu()[same], ref["photometric_confidence"][same], rtol=conf_tol, atol=conf_tol)
    return mismatch


def test_cfg1_full_against_oracle():
    """BASELINE cfg1 in full: 640x512, N=3, D=48 inverse-depth planes, model_dtu_v2.ckpt, against the oracle run on the box's
    host cores (~25 s)."""
    from aarmvsnet_b200 import synth
    from oracle import aarmvs_oracle as orc
    s = synth.make_sample(512, 640, 3, 48, seed=33, depth_mode="inverse")
    out = _run(_model(512, 640), s)
    ref = orc.forward_test(load_weights("model_dtu_v2.ckpt"), s["imgs"], s["proj_matrices"], s["depth_values"], return_index=True,
                           return_cost=True)
    _compare("cfg1 640x512 N=3 D=48", out, ref, s)


def test_long_recurrence_d512_against_oracle():
    """The headline depth count: D = 512 dependent ConvLSTM steps (split-fp16 products, approximate gate transcendentals and
    fp64-vs-ATen GroupNorm sums accumulate through the recurrence) at 128x160, N=5, uniform depths, against the oracle; also
    prints the drift of the regularised cost at planes 1 / 64 / 256 / 511."""
    from aarmvsnet_b200 import synth
    from oracle import aarmvs_oracle as orc
    H, W, N, D = 128, 160, 5, 512
    s = synth.make_sample(H, W, N, D, seed=41)
    m = _model(H, W)
    out = _run(m, s)
    ref = orc.forward_test(load_weights("model_dtu_v2.ckpt"), s["imgs"], s["proj_matrices"], s["depth_values"], return_index=True,
                           return_cost=True)
    o = m._engine.depth_map(s["imgs"][0].cuda(), s["proj_matrices"][0], s["depth_values"][0].cuda(), keep_cost=True)
    cost = o["cost_reg"].cpu()
    drift = {d: (cost[d] - ref["cost_reg"][0, d]).abs().max().item() for d in (1, 64, 256, 511)}
    print("D=512 recurrence: max |cost_reg - oracle| at planes 1/64/256/511:", ", ".join(f"{v:.2e}" for v in drift.values()))
    assert max(drift.values()) < 2e-3
    _compare("128x160 N=5 D=512", out, ref, s)


def test_blended_mid_size_against_oracle():
    """model_blended_v2.ckpt beyond the 40x56 golden fixture: 128x160, N=7, D=64 inverse depths."""
    from aarmvsnet_b200 import synth
    from oracle import aarmvs_oracle as orc
    s = synth.make_sample(128, 160, 7, 64, seed=43, depth_mode="inverse")
    out = _run(_model(128, 160, ckpt="model_blended_v2.ckpt"), s)
    ref = orc.forward_test(load_weights("model_blended_v2.ckpt"), s["imgs"], s["proj_matrices"], s["depth_values"], return_index=True,
                           return_cost=True)
    _compare("blended 128x160 N=7 D=64", out, ref, s)


def test_tensor_core_and_fma_cells_agree():
    """The split-fp16 tcgen05 path against the fp32 CUDA-core path on the same inputs (no CPU needed): 320x256, D=24."""
    from aarmvsnet_b200 import synth
    s = synth.make_sample(256, 320, 4, 24, seed=5)
    a = _run(_model(256, 320, cell_math="tc"), s)
    b = _run(_model(256, 320, cell_math="fma"), s)
    same = (a["index"] == b["index"]).float().mean().item()
    print(f"tc vs fma: same plane on {same * 100:.4f}% of pixels")
    assert same >= 1.0 - MISMATCH_BAR
    assert (a["photometric_confidence"] - b["photometric_confidence"]).abs().max().item() < 5e-3


def test_plane_batching_does_not_change_the_result():
    from aarmvsnet_b200 import synth
    s = synth.make_sample(128, 160, 4, 19, seed=9)          # 19 planes: batches of 8, 8, 3
    a = _run(_model(128, 160, plane_batch=8), s)
    b = _run(_model(128, 160, plane_batch=1), s)
    assert (a["index"] == b["index"]).float().mean().item() >= 0.9999
    torch.testing.assert_close(a["photometric_confidence"], b["photometric_confidence"], rtol=1e-4, atol=1e-5)


def test_cfg2_full_size_properties():
    """BASELINE cfg2.  Properties that hold for any correct forward: the depth is one of the hypotheses and matches the
    winning index, 0 < confidence <= 1, everything finite, and a second call (graph replay) reproduces the first."""
    from aarmvsnet_b200 import synth
    H, W, N, D = 1184, 1600, 7, 512
    s = synth.make_sample(H, W, N, D, seed=3)
    m = _model(H, W)
    a = _run(m, s)
    dv = s["depth_values"][0].cuda()
    idx = a["index"][0].long()
    assert idx.min().item() >= 0 and idx.max().item() < D
    assert torch.equal(a["depth"][0], dv[idx])
    conf = a["photometric_confidence"][0]
    assert torch.isfinite(conf).all() and conf.min().item() > 0 and conf.max().item() <= 1.0
    assert idx.float().std().item() > 1.0                     # not a degenerate constant map
    b = _run(m, s)
    assert (a["index"] == b["index"]).float().mean().item() >= 0.99999
    torch.testing.assert_close(a["photometric_confidence"], b["photometric_confidence"], rtol=1e-4, atol=1e-5)


def test_cfg4_full_size_properties():
    """BASELINE cfg4 (Tanks&Temples shape 1920x1056, N=7, D=512, model_blended_v2.ckpt, inverse depths): the same
    size-independent properties as cfg2."""
    from aarmvsnet_b200 import synth
    H, W, N, D = 1056, 1920, 7, 512
    s = synth.make_sample(H, W, N, D, seed=4, depth_mode="inverse")
    a = _run(_model(H, W, ckpt="model_blended_v2.ckpt"), s)
    dv = s["depth_values"][0].cuda()
    idx = a["index"][0].long()
    assert idx.min().item() >= 0 and idx.max().item() < D
    assert torch.equal(a["depth"][0], dv[idx])
    conf = a["photometric_confidence"][0]
    assert torch.isfinite(conf).all() and conf.min().item() > 0 and conf.max().item() <= 1.0
    assert idx.float().std().item() > 1.0


def test_cfg2_full_resolution_few_planes_against_oracle():
    """BASELINE cfg2's image size and view count (1600x1184, N=7) with D=6 planes, against the CPU oracle run on the box's host
    cores (~40 s): the tiling, the 32-bit byte offsets and the ragged last tiles of every kernel at the real resolution, where
    the full D=512 run is only property-checked."""
    from aarmvsnet_b200 import synth
    from oracle import aarmvs_oracle as orc
    H, W, N, D = 1184, 1600, 7, 6
    s = synth.make_sample(H, W, N, D, seed=5, smooth=True)
    out = _run(_model(H, W), s)
    ref = orc.forward_test(load_weights("model_dtu_v2.ckpt"), s["imgs"], s["proj_matrices"], s["depth_values"], return_index=True)
    idx = out["index"].cpu().long()
    same = idx == ref["index"]
    mismatch = 1.0 - same.float().mean().item()
    print(f"cfg2 resolution, D=6: plane mismatch fraction {mismatch:.2e} on {same.numel()} pixels")
    assert mismatch <= MISMATCH_BAR
    assert (out["depth"].cpu()[same] == ref["depth"][same]).all()
    torch.testing.assert_close(out["photometric_confidence"].cpu()[same], ref["photometric_confidence"][same], rtol=1e-3, atol=1e-3)


def test_cfg4_full_resolution_few_planes_against_oracle():
    """BASELINE cfg4's image size, view count and weights (1920x1056, N=7, model_blended_v2.ckpt, inverse depths) with D=4 planes
    against the CPU oracle on the box's host cores (~45 s): the full D=512 run is only property-checked."""
    from aarmvsnet_b200 import synth
    from oracle import aarmvs_oracle as orc
    H, W, N, D = 1056, 1920, 7, 4
    s = synth.make_sample(H, W, N, D, seed=6, depth_mode="inverse", smooth=True)
    out = _run(_model(H, W, ckpt="model_blended_v2.ckpt"), s)
    ref = orc.forward_test(load_weights("model_blended_v2.ckpt"), s["imgs"], s["proj_matrices"], s["depth_values"], return_index=True,
                           return_cost=True)
    _compare("cfg4 resolution 1920x1056 N=7 D=4 (blended)", out, ref, s)
