"""Pins oracle/aarmvs_oracle.py against outputs of the unmodified reference (tests/golden, made by
oracle/make_golden.py).  The restatement uses the same ATen ops, so agreement is expected to ~1 ulp."""
import pytest
import torch

from conftest import load_golden, load_weights
from oracle import aarmvs_oracle as orc

TOL = dict(rtol=1e-5, atol=1e-6)


def test_featnet(ops_dtu, sd_dtu):
    out = orc.feat_net(ops_dtu["feat_img"], sd_dtu)
    torch.testing.assert_close(out, ops_dtu["feat_out"], **TOL)


def test_deform_conv(ops_dtu, sd_dtu):
    out = orc.deform_conv(ops_dtu["deform_in"], sd_dtu, "feature.intraAA.deformconv1.0")
    torch.testing.assert_close(out, ops_dtu["deform_out"], rtol=1e-4, atol=1e-5)
    big = orc.deform_conv(ops_dtu["deform_big_in"], sd_dtu, "feature.intraAA.deformconv0.0")
    torch.testing.assert_close(big, ops_dtu["deform_big_out"], rtol=1e-4, atol=1e-4)


def test_warp(ops_dtu):
    g = ops_dtu
    out = orc.homo_warp(g["warp_src"], g["warp_src_proj"], g["warp_ref_proj"], g["warp_depth"])
    torch.testing.assert_close(out, g["warp_out"], rtol=0, atol=0)
    far = orc.homo_warp(g["warp_src"], g["warp_far_proj"], g["warp_ref_proj"], g["warp_depth"])
    torch.testing.assert_close(far, g["warp_far_out"], rtol=0, atol=0)
    assert (g["warp_far_out"] == 0).float().mean() > 0.1     # the far camera really leaves the image


def test_omega(ops_dtu, sd_dtu):
    torch.testing.assert_close(orc.omega(ops_dtu["omega_in"], sd_dtu), ops_dtu["omega_out"], **TOL)


def test_cell(ops_dtu, sd_dtu):
    g = ops_dtu
    h, c = orc.convlstm_cell(g["cell_x"], g["cell_h"], g["cell_c"], sd_dtu, "cost_regularization.cell_list.0")
    torch.testing.assert_close(h, g["cell_h_out"], **TOL)
    torch.testing.assert_close(c, g["cell_c_out"], **TOL)


def test_unet(ops_dtu, sd_dtu):
    g = ops_dtu
    state = orc.init_state(1, 32, 40)
    for i in range(g["unet_x"].shape[0]):
        cost, state = orc.unet_step(g["unet_x"][i], state, sd_dtu)
        torch.testing.assert_close(cost, g["unet_cost"][i], **TOL)
    for (h, c), (gh, gc) in zip(state, g["unet_state"]):
        torch.testing.assert_close(h, gh, **TOL)
        torch.testing.assert_close(c, gc, **TOL)


@pytest.mark.parametrize("name", ["e2e_48x64_n3_d8.pt", "e2e_64x80_n5_d24_smooth.pt",
                                  "e2e_blended_40x56_n4_d12_b2.pt"])
def test_forward_test_phase(name):
    g = load_golden(name)
    sd = load_weights(g["ckpt"])
    out = orc.forward_test(sd, g["imgs"], g["proj_matrices"], g["depth_values"], return_index=True,
                           return_cost=True)
    torch.testing.assert_close(out["cost_reg"], g["cost_reg"], rtol=1e-4, atol=1e-4)
    assert (out["index"] == g["index"].long()).float().mean() >= 0.9995
    same = out["index"] == g["index"].long()
    torch.testing.assert_close(out["depth"][same], g["depth"][same], rtol=0, atol=0)
    torch.testing.assert_close(out["photometric_confidence"][same], g["photometric_confidence"][same],
                               rtol=1e-4, atol=1e-5)


def test_forward_train_phase_and_loss():
    g = load_golden("train_32x40_n3_d6.pt")
    sd = load_weights(g["ckpt"])
    prob = orc.forward_train(sd, g["imgs"], g["proj_matrices"], g["depth_values"])
    torch.testing.assert_close(prob, g["prob_volume"], rtol=1e-4, atol=1e-6)
    loss, wta = orc.cls_loss(g["prob_volume"], g["depth_gt"], g["mask"], g["depth_values"])
    torch.testing.assert_close(loss, g["loss"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(wta, g["wta_depth"], rtol=0, atol=0)


def test_training_gradients():
    """Backward of the training step (train.py:288-300): autograd through the oracle's train-phase forward and loss must
    reproduce d loss / d parameter of the UNMODIFIED reference for all 90 parameters (tests/golden/train_grads_*.pt,
    oracle/make_golden.py grads).  This pins the reference the backward kernels of config 5 are to be checked against."""
    g = load_golden("train_32x40_n3_d6.pt")
    gg = load_golden("train_grads_32x40_n3_d6.pt")
    sd = {k: v.clone().requires_grad_(True) for k, v in load_weights(g["ckpt"]).items()}
    prob = orc.forward_train(sd, g["imgs"], g["proj_matrices"], g["depth_values"])
    loss, _ = orc.cls_loss(prob, g["depth_gt"], g["mask"], g["depth_values"])
    torch.testing.assert_close(loss.detach(), gg["loss"], rtol=1e-5, atol=1e-6)
    loss.backward()
    assert set(gg["grads"]) == set(sd) and len(sd) == 90
    worst = 0.0
    top = max(v.abs().max().item() for v in gg["grads"].values())
    for k, want in gg["grads"].items():
        got = sd[k].grad
        assert got is not None, k
        # conv_0.bias shifts every plane's cost alike, so its true gradient through the softmax is 0 (1.5e-8 of rounding noise
        # in the reference): tensors are compared on the scale of their own largest gradient, with a floor from the global one
        scale = max(want.abs().max().item(), 1e-4 * top)
        worst = max(worst, (got - want).abs().max().item() / scale)
        torch.testing.assert_close(got, want, rtol=1e-3, atol=2e-4 * scale, msg=lambda m, k=k: f"{k}: {m}")
    print(f"worst max-abs gradient error relative to the tensor's largest gradient: {worst:.2e}")
