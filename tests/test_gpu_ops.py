"""GPU parity, operator by operator: the CUDA kernels (through the C ABI) against the golden outputs of the
unmodified reference (tests/golden/ops_dtu.pt) and against the CPU oracle on the same inputs."""
import pytest
import torch

from conftest import load_weights
from oracle import aarmvs_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    import aarmvsnet_b200 as aa
    m = aa.EMVSNet(disparity_level=32, image_scale=1.0, max_h=32, max_w=40, return_depth=True)
    m.load_state_dict(load_weights("model_dtu_v2.ckpt"), strict=True)
    return m.cuda().eval()


def test_warp_matches_reference_bit_for_bit(ops_dtu):
    """homo_warping_depthwise: the kernel mirrors ATen's fp32 op order, so almost every element is bit-identical."""
    from aarmvsnet_b200 import homo_warping_depthwise
    g = ops_dtu
    for proj, want in ((g["warp_src_proj"], g["warp_out"]), (g["warp_far_proj"], g["warp_far_out"])):
        got = homo_warping_depthwise(g["warp_src"].cuda(), proj.cuda(), g["warp_ref_proj"].cuda(), g["warp_depth"].cuda()).cpu()
        assert (got == want).float().mean() >= 0.999
        torch.testing.assert_close(got, want, rtol=0, atol=2e-6)
        assert ((got == 0) == (want == 0)).all()          # same out-of-image pattern (zeros padding per tap)


def test_warp_align_corners_switch(ops_dtu):
    import torch.nn.functional as F
    from aarmvsnet_b200 import homo_warping_depthwise
    g = ops_dtu
    got = homo_warping_depthwise(g["warp_src"].cuda(), g["warp_src_proj"].cuda(), g["warp_ref_proj"].cuda(),
                                 g["warp_depth"].cuda(), align_corners=True).cpu()
    # oracle with the old (torch <= 1.2) grid_sample convention
    b, c, h, w = g["warp_src"].shape
    rot, trans = orc.relative_projection(g["warp_src_proj"], g["warp_ref_proj"])
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    xyz = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(h * w))).unsqueeze(0)
    p = torch.matmul(rot, xyz) * g["warp_depth"].view(b, 1, 1) + trans
    xy = p[:, :2] / p[:, 2:3]
    grid = torch.stack((xy[:, 0] / ((w - 1) / 2) - 1, xy[:, 1] / ((h - 1) / 2) - 1), dim=2).view(b, h, w, 2)
    want = F.grid_sample(g["warp_src"], grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    torch.testing.assert_close(got, want, rtol=0, atol=1e-5)


def test_featnet(ops_dtu, model):
    from aarmvsnet_b200 import ops
    got = ops.feat_net(model, ops_dtu["feat_img"].cuda()).cpu()
    torch.testing.assert_close(got, ops_dtu["feat_out"], rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("math", ["tc", "fma"])
def test_deform_conv(ops_dtu, model, math):
    """tc: gathered operand staged as split fp16 for tcgen05; fma: fp32 CUDA-core kernel.  Same tolerance for both."""
    from aarmvsnet_b200 import ops
    got = ops.deform_conv(model, 1, ops_dtu["deform_in"][0].cuda(), math=math).cpu()
    torch.testing.assert_close(got, ops_dtu["deform_out"][0], rtol=1e-4, atol=1e-4)
    # large offsets: sampling positions clamp at the padded border (module.py:153-159)
    big = ops.deform_conv(model, 0, ops_dtu["deform_big_in"][0].cuda(), math=math).cpu()
    torch.testing.assert_close(big, ops_dtu["deform_big_out"][0], rtol=1e-4, atol=5e-4)


@pytest.mark.parametrize("math", ["tc", "fma"])
def test_convlstm_cell(ops_dtu, model, math):
    """tc: tcgen05 implicit GEMM with split-fp16 operands; fma: fp32 CUDA-core kernel.  Same tolerance for both."""
    from aarmvsnet_b200 import ops
    g = ops_dtu
    h, c = ops.convlstm_cell(model, 0, g["cell_x"][0].cuda(), g["cell_h"][0].cuda(), g["cell_c"][0].cuda(), math=math)
    torch.testing.assert_close(h.cpu(), g["cell_h_out"][0], rtol=1e-5, atol=2e-5)
    torch.testing.assert_close(c.cpu(), g["cell_c_out"][0], rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("hw", [(16, 8), (20, 28), (37, 50)])
def test_deconv_gn_relu_tensor_core(sd_dtu, model, hw):
    """deConvGnReLU as a two-pass tcgen05 GEMM (statistics, then normalise + split) against torch's ConvTranspose2d +
    GroupNorm + ReLU, incl. sizes that do not fill the 16x8 tiles."""
    import torch.nn.functional as F
    from aarmvsnet_b200 import ops
    h, w = hw
    x = torch.randn(16, h, w, generator=torch.Generator().manual_seed(h * 100 + w))
    for i in range(2):
        p = f"cost_regularization.deconv_{i}"
        y = F.conv_transpose2d(x[None], sd_dtu[p + ".conv.weight"], sd_dtu[p + ".conv.bias"], stride=2, padding=1, output_padding=1)
        want = F.relu(F.group_norm(y, 2, sd_dtu[p + ".gn.weight"], sd_dtu[p + ".gn.bias"], eps=1e-5))[0]
        got = ops.deconv_gn_relu(model, i, x.cuda()).cpu()
        torch.testing.assert_close(got, want, rtol=1e-5, atol=2e-5)


@pytest.fixture(scope="module")
def model_fma():
    import aarmvsnet_b200 as aa
    m = aa.EMVSNet(disparity_level=32, image_scale=1.0, max_h=32, max_w=40, return_depth=True, cell_math="fma")
    m.load_state_dict(load_weights("model_dtu_v2.ckpt"), strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("math", ["tc", "tc-gather", "fma"])
@pytest.mark.parametrize("hw", [(32, 40), (36, 52), (20, 100)])
def test_aggregate_plane_and_omega(ops_dtu, sd_dtu, model, model_fma, hw, math):
    """tc: pass A keeps (warp-ref)^2 and pass D streams it; tc-gather / fma: pass D gathers a second time."""
    stream_costv = math == "tc"
    model = model_fma if math == "fma" else model
    """warp + (warp-ref)^2 + omega + weighted sum, against the oracle, incl. sizes that do not fill the 16x32 tiles."""
    from aarmvsnet_b200 import ops, synth
    h, w = hw
    gen = torch.Generator().manual_seed(h * 1000 + w)
    s = synth.make_sample(h, w, 4, 4, seed=h)
    feats = [torch.randn(1, 32, h, w, generator=gen) for _ in range(4)]
    pm, dv = s["proj_matrices"], s["depth_values"][:, 1]
    want = orc.aggregate_plane(feats[0], feats[1:], pm[:, 0], [pm[:, v] for v in range(1, 4)], dv, sd_dtu)
    got, rew = ops.aggregate_plane(model, feats[0][0].cuda(), [f[0].cuda() for f in feats[1:]], pm[0, 0],
                                   [pm[0, v] for v in range(1, 4)], dv[0], stream_costv=stream_costv)
    for v in range(3):
        c = (orc.homo_warp(feats[v + 1], pm[:, v + 1], pm[:, 0], dv) - feats[0]).pow(2)
        torch.testing.assert_close(rew[v].cpu(), orc.omega(c, sd_dtu)[0, 0], rtol=1e-4, atol=2e-5)
    torch.testing.assert_close(got.cpu(), want[0], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_same_process_drives_two_devices(sd_dtu):
    """One process, two GPUs: launch attributes (dynamic shared memory opt-in) and the SM count are per-device properties
    (a process-wide flag made every large-smem conv fail on the second device)."""
    import aarmvsnet_b200 as aa
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    s = aa.synth.make_sample(48, 64, 3, 6, seed=5)
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        m = aa.EMVSNet(32, image_scale=1.0, max_h=48, max_w=64, return_depth=True, cell_math="tc" if dev.endswith("0") else "fma")
        m.load_state_dict(sd_dtu, strict=True)
        m = m.to(dev).eval()
        o = m(s["imgs"].to(dev), s["proj_matrices"].to(dev), s["depth_values"].to(dev))
        outs.append(o["index"].cpu())
        m2 = aa.EMVSNet(32, image_scale=1.0, max_h=48, max_w=64, return_depth=True, cell_math="fma" if dev.endswith("0") else "tc")
        m2.load_state_dict(sd_dtu, strict=True)
        outs.append(m2.to(dev).eval()(s["imgs"].to(dev), s["proj_matrices"].to(dev), s["depth_values"].to(dev))["index"].cpu())
    for o in outs[1:]:
        assert (o == outs[0]).float().mean().item() >= 0.999
