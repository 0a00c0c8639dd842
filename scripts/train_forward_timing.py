"""Where the forward + loss part of a cfg5 training step goes: the engine's per-sample forwards (concurrent sweeps) against the
torch ops on top (stack, softmax, mvsnet_cls_loss), device-timed.
    python scripts/train_forward_timing.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aarmvsnet_b200 as aa  # noqa: E402
from aarmvsnet_b200.training import mvsnet_cls_loss  # noqa: E402

H, W, N, D, B = 128, 160, 3, 192, 4
sd = torch.load(os.path.join(ROOT, "tests", "golden", "model_dtu_v2.ckpt"), map_location="cpu", weights_only=False)["model"]
m = aa.EMVSNet(32, image_scale=1.0, max_h=H, max_w=W, return_depth=False)
m.load_state_dict(sd, strict=True)
m = m.cuda().train()
eng = m._engine
ss = [aa.synth.make_sample(H, W, N, D, seed=3 + b, depth_mode="inverse") for b in range(B)]
imgs = torch.cat([s["imgs"] for s in ss]).cuda()
proj = torch.cat([s["proj_matrices"] for s in ss])
dv = torch.cat([s["depth_values"] for s in ss]).cuda()
gt = torch.rand(B, H, W, device="cuda") * 500 + 400
mask = torch.ones(B, H, W, device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
for rep in range(3):
    main = torch.cuda.current_stream()
    streams = eng.sweep_streams(imgs.device, B)
    torch.cuda.synchronize()
    ev[0].record()
    saved = [eng.forward_saving_state(imgs[b], proj[b], dv[b], slot=b, sweep_stream=streams[b]) for b in range(B)]
    ev[1].record()                                   # FeatNet of all samples done on the main stream
    for st in streams:
        main.wait_stream(st)
    ev[2].record()
    cost = torch.stack([s["cost"].view(-1, H, W) for s in saved]).requires_grad_(True)
    prob = torch.softmax(cost, dim=1)
    ev[3].record()
    loss, depth = mvsnet_cls_loss(prob, gt, mask, dv)
    ev[4].record()
    (dcost,) = torch.autograd.grad(loss, cost)
    ev[5].record()
    torch.cuda.synchronize()
    print(f"FeatNet x{B} (main stream) {ev[0].elapsed_time(ev[1]):.1f} ms | + sweeps joined {ev[0].elapsed_time(ev[2]):.1f} | stack+softmax "
          f"{ev[2].elapsed_time(ev[3]):.1f} | loss {ev[3].elapsed_time(ev[4]):.1f} | loss backward to dcost {ev[4].elapsed_time(ev[5]):.1f}")
