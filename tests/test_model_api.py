"""Host-side contract of the drop-in (no GPU): constructor / state-dict compatibility with the shipped checkpoints,
weight repacking, error behaviour, synthetic-sample generator."""
import pytest
import torch

from conftest import load_weights


@pytest.mark.parametrize("ckpt", ["model_dtu_v2.ckpt", "model_blended_v2.ckpt"])
def test_shipped_checkpoints_load_strict(ckpt):
    import aarmvsnet_b200 as aa
    sd = load_weights(ckpt)
    m = aa.EMVSNet(disparity_level=32, image_scale=1.0, max_h=512, max_w=640, return_depth=True)
    assert len(sd) == 90 and sum(v.numel() for v in sd.values()) == 187203
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    got = m.state_dict()
    assert list(got.keys()) != [] and set(got.keys()) == set(sd.keys())
    for k in sd:
        assert got[k].shape == sd[k].shape, k
    # 'module.' prefixed checkpoints (nn.DataParallel, eval.py:93-95) load after stripping like the reference does
    m.load_state_dict({k[len("module."):]: v for k, v in {"module." + k: v for k, v in sd.items()}.items()}, strict=True)


def test_reference_constructor_signature():
    import inspect
    import aarmvsnet_b200 as aa
    p = inspect.signature(aa.EMVSNet.__init__).parameters
    assert list(p)[1:6] == ["disparity_level", "image_scale", "max_h", "max_w", "return_depth"]   # drmvsnet.py:235
    assert p["image_scale"].default == 0.25 and p["max_h"].default == 960 and p["max_w"].default == 480
    assert p["return_depth"].default is False


def test_no_cpu_path():
    import aarmvsnet_b200 as aa
    m = aa.EMVSNet(32, image_scale=1.0, max_h=16, max_w=16, return_depth=True)
    s = aa.synth.make_sample(16, 16, 3, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(s["imgs"], s["proj_matrices"], s["depth_values"])
    with pytest.raises(AssertionError):
        m(s["imgs"], s["proj_matrices"][:, :2], s["depth_values"])


def test_weight_repacking():
    from aarmvsnet_b200 import pack
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).view(2, 3, 3, 3)
    t = pack.conv_taps(w, cin_pad=4, cout_pad=4)
    assert t.shape == (9, 4, 4)
    assert t[5, 1, 1] == w[1, 1, 1, 2] and t[5, 3, 0] == 0 and t[0, 0, 3] == 0
    wl = torch.randn(8, 5, 3, 3)
    bl = torch.arange(8, dtype=torch.float32)
    pw, pb = pack.lstm_gates(wl, bl)           # hidden = 2: [i0 i1 f0 f1 o0 o1 g0 g1] -> [i0 f0 o0 g0 i1 f1 o1 g1]
    assert pb.tolist() == [0, 2, 4, 6, 1, 3, 5, 7]
    assert torch.equal(pw[4, :, 1], wl[2, :, 1, 1])
    wt = torch.randn(16, 16, 3, 3)
    assert torch.equal(pack.deconv_taps(wt)[2, 0, 3, 5], wt[3, 5, 2, 0])
    sd = load_weights("model_dtu_v2.ckpt")
    tail = pack.omega_tail(sd)
    assert tail.numel() == 69
    assert torch.equal(tail[8:24].view(4, 4), sd["omega.reweight_network.1.stem.0.0.weight"].view(4, 4))


def test_synth_sample_is_deterministic_and_valid():
    from aarmvsnet_b200 import synth
    a = synth.make_sample(32, 40, 4, 6, seed=3)
    b = synth.make_sample(32, 40, 4, 6, seed=3)
    for k in a:
        assert torch.equal(a[k], b[k])
    pm = a["proj_matrices"]
    assert pm.shape == (1, 4, 4, 4) and torch.equal(pm[0, :, 3], torch.tensor([[0., 0, 0, 1]]).repeat(4, 1))
    # the reference frustum's centre projects inside every source image
    x = pm[0] @ torch.tensor([0.0, 0.0, 680.0, 1.0])
    u, v = x[:, 0] / x[:, 2], x[:, 1] / x[:, 2]
    assert ((u > 0) & (u < 40) & (v > 0) & (v < 32)).all()
    dv = synth.depth_hypotheses(8, "inverse")
    assert torch.all(dv[1:] > dv[:-1]) and abs(dv[0].item() - 425) < 1e-3


def test_cls_loss_matches_reference_fixture():
    """mvsnet_cls_loss (models/drmvsnet.py:347-381) against the value the unmodified reference produced."""
    import aarmvsnet_b200 as aa
    from conftest import load_golden
    g = load_golden("train_32x40_n3_d6.pt")
    loss, wta = aa.mvsnet_cls_loss(g["prob_volume"], g["depth_gt"], g["mask"], g["depth_values"])
    torch.testing.assert_close(loss, g["loss"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(wta, g["wta_depth"], rtol=0, atol=0)
    loss2, _, conf = aa.mvsnet_cls_loss(g["prob_volume"], g["depth_gt"], g["mask"], g["depth_values"], return_prob_map=True)
    assert torch.equal(loss2, loss) and torch.equal(conf, g["prob_volume"].max(1)[0])
