"""Generate tests/golden/*.pt by running the UNMODIFIED reference (/root/reference) on CPU.

Runs only in the build container (the GPU box has no /root/reference).  Recipe = SURVEY.md §8c:
``sys.path.append`` (never insert(0): /root/reference/statistics.py shadows the stdlib module), ``.cuda()``
patched to identity (module.py:94-96 and drmvsnet.py:302-304 hard-code it), evidential head stubbed (no
weights in the shipped checkpoints, raises for D != 32), ``strict=False`` load, ``eval()``, ``no_grad()``.

    python oracle/make_golden.py            # rewrites tests/golden/*.pt

The fixtures pin oracle/aarmvs_oracle.py (tests/test_oracle_golden.py) and are the known answers the CUDA
path is compared with on the GPU box (tests/test_gpu_*.py).
"""
import os
import shutil
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)
sys.path.append(REF)

torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self

from models import module as ref_module          # noqa: E402
from models.drmvsnet import EMVSNet              # noqa: E402
from aarmvsnet_b200 import synth                 # noqa: E402


class _NoEvidential(torch.nn.Module):
    def forward(self, prob_volume, depth_values):
        return torch.zeros(1), None


def load_reference(ckpt, h, w, return_depth=True):
    m = EMVSNet(disparity_level=32, image_scale=1.0, max_h=h, max_w=w, return_depth=return_depth)
    sd = torch.load(os.path.join(REF, "checkpoints", ckpt), map_location="cpu", weights_only=False)["model"]
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.startswith("evidential.") for k in res.missing_keys)
    m.evidential = _NoEvidential()
    m.eval()
    return m


def ops_fixture(ckpt, seed=7):
    g = torch.Generator().manual_seed(seed)
    h, w = 32, 40
    m = load_reference(ckpt, h, w)
    out = {"ckpt": ckpt}
    with torch.no_grad():
        img = torch.randn(1, 3, h, w, generator=g)
        out["feat_img"] = img
        out["feat_out"] = m.feature(img)
        # intermediate trunk activations for per-layer checks
        x = m.feature.init_conv(img)
        x0 = m.feature.conv0(x)
        x1 = m.feature.conv1(x0)
        x2 = m.feature.conv2(x1)
        out["feat_x0"], out["feat_x1"], out["feat_x2"] = x0, x1, x2
        dc = m.feature.intraAA.deformconv1[0]
        out["deform_in"] = x1
        out["deform_out"] = dc(x1)
        # a second deformable case with large synthetic offsets so clamping at the border is exercised
        big = torch.randn(1, 32, 12, 16, generator=g) * 3.0
        out["deform_big_in"] = big
        out["deform_big_out"] = m.feature.intraAA.deformconv0[0](big)

        sample = synth.make_sample(h, w, 3, 4, seed=seed)
        projs = sample["proj_matrices"]
        src = torch.randn(1, 32, h, w, generator=g)
        ref = torch.randn(1, 32, h, w, generator=g)
        dv = sample["depth_values"][:, 2]
        out["warp_src"], out["warp_ref"] = src, ref
        out["warp_src_proj"], out["warp_ref_proj"], out["warp_depth"] = projs[:, 1], projs[:, 0], dv
        warped = ref_module.homo_warping_depthwise(src, projs[:, 1], projs[:, 0], dv)
        out["warp_out"] = warped
        # a far-off camera: most samples leave the image (zeros padding, partial taps)
        far = projs[:, 2].clone()
        far[:, 0, 3] += 4000.0
        out["warp_far_proj"] = far
        out["warp_far_out"] = ref_module.homo_warping_depthwise(src, far, projs[:, 0], dv)
        cost_v = (warped - ref).pow(2)
        out["omega_in"] = cost_v
        out["omega_out"] = m.omega(cost_v)

        cell = m.cost_regularization.cell_list[0]
        cx = torch.randn(1, 32, h, w, generator=g)
        ch = torch.randn(1, 16, h, w, generator=g) * 0.5
        cc = torch.randn(1, 16, h, w, generator=g)
        out["cell_x"], out["cell_h"], out["cell_c"] = cx, ch, cc
        out["cell_h_out"], out["cell_c_out"] = cell(cx, (ch, cc))

        xs = [-(torch.randn(1, 32, h, w, generator=g) ** 2) for _ in range(3)]
        out["unet_x"] = torch.stack(xs, 0)
        state, costs = None, []
        for i, xi in enumerate(xs):
            cost, state = m.cost_regularization(xi, state, i)
            costs.append(cost)
        out["unet_cost"] = torch.stack(costs, 0)
        out["unet_state"] = [(hh.clone(), cc_.clone()) for hh, cc_ in state]
    return out


def e2e_fixture(ckpt, h, w, n, d, seed, depth_mode="uniform", smooth=False, batch=1):
    m = load_reference(ckpt, h, w)
    sample = synth.make_sample(h, w, n, d, seed=seed, batch=batch, depth_mode=depth_mode, smooth=smooth)
    costs = []
    hook = m.cost_regularization.register_forward_hook(lambda mod, inp, outp: costs.append(outp[0].squeeze(1).clone()))
    with torch.no_grad():
        res = m(sample["imgs"], sample["proj_matrices"], sample["depth_values"])
    hook.remove()
    dv = sample["depth_values"]
    index = (res["depth"].unsqueeze(1) - dv.view(batch, d, 1, 1)).abs().argmin(1)
    return {"ckpt": ckpt, "h": h, "w": w, "n": n, "d": d, "seed": seed, "depth_mode": depth_mode, "smooth": smooth,
            "batch": batch, **sample, "depth": res["depth"], "photometric_confidence": res["photometric_confidence"],
            "index": index.to(torch.int16), "cost_reg": torch.stack(costs, 1)}


def train_fixture(ckpt, h, w, n, d, seed):
    from models.drmvsnet import mvsnet_cls_loss
    m = load_reference(ckpt, h, w, return_depth=False)
    sample = synth.make_sample(h, w, n, d, seed=seed, depth_mode="inverse")
    with torch.no_grad():
        prob, _, _ = m(sample["imgs"], sample["proj_matrices"], sample["depth_values"])
        g = torch.Generator().manual_seed(seed + 1)
        depth_gt = 425.0 + 510.0 * torch.rand(1, h, w, generator=g)
        mask = (torch.rand(1, h, w, generator=g) > 0.2).float()
        loss, wta = mvsnet_cls_loss(prob, depth_gt, mask, sample["depth_values"])
    return {"ckpt": ckpt, "h": h, "w": w, "n": n, "d": d, "seed": seed, **sample, "prob_volume": prob,
            "depth_gt": depth_gt, "mask": mask, "loss": loss, "wta_depth": wta}


def grads_fixture(train_path):
    """Backward of the training step (train.py:288-300: loss = mvsnet_cls_loss(model(...)); loss.backward()) of the UNMODIFIED
    reference on the sample of the train fixture: d loss / d parameter for all 90 parameters.  Pins the autograd of the oracle
    (tests/test_oracle_golden.py::test_training_gradients), the reference the round-2 backward kernels are checked against."""
    from models.drmvsnet import mvsnet_cls_loss
    t = torch.load(train_path, map_location="cpu", weights_only=False)
    m = load_reference(t["ckpt"], t["h"], t["w"], return_depth=False)
    for p in m.parameters():
        p.requires_grad_(True)
    prob, _, _ = m(t["imgs"], t["proj_matrices"], t["depth_values"])
    loss, _ = mvsnet_cls_loss(prob, t["depth_gt"], t["mask"], t["depth_values"])
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    assert abs(loss.item() - float(t["loss"])) <= 1e-5 * abs(float(t["loss"]))
    return {"train_fixture": os.path.basename(train_path), "loss": loss.detach(), "grads": grads}


def main():
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "grads":          # only the gradient fixture (the others stay untouched)
        g = grads_fixture(os.path.join(GOLD, "train_32x40_n3_d6.pt"))
        torch.save(g, os.path.join(GOLD, "train_grads_32x40_n3_d6.pt"))
        print("train_grads_32x40_n3_d6.pt", len(g["grads"]), "parameter gradients, loss", float(g["loss"]))
        return
    for ck in ("model_dtu_v2.ckpt", "model_blended_v2.ckpt"):
        dst = os.path.join(GOLD, ck)
        if not os.path.exists(dst):
            shutil.copyfile(os.path.join(REF, "checkpoints", ck), dst)   # shipped weights (data, not source)
    torch.save(ops_fixture("model_dtu_v2.ckpt"), os.path.join(GOLD, "ops_dtu.pt"))
    torch.save(e2e_fixture("model_dtu_v2.ckpt", 48, 64, 3, 8, seed=1), os.path.join(GOLD, "e2e_48x64_n3_d8.pt"))
    torch.save(e2e_fixture("model_dtu_v2.ckpt", 64, 80, 5, 24, seed=2, depth_mode="inverse", smooth=True),
               os.path.join(GOLD, "e2e_64x80_n5_d24_smooth.pt"))
    torch.save(e2e_fixture("model_blended_v2.ckpt", 40, 56, 4, 12, seed=3, depth_mode="inverse", batch=2),
               os.path.join(GOLD, "e2e_blended_40x56_n4_d12_b2.pt"))
    torch.save(train_fixture("model_dtu_v2.ckpt", 32, 40, 3, 6, seed=4), os.path.join(GOLD, "train_32x40_n3_d6.pt"))
    torch.save(grads_fixture(os.path.join(GOLD, "train_32x40_n3_d6.pt")), os.path.join(GOLD, "train_grads_32x40_n3_d6.pt"))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
