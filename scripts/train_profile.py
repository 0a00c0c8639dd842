"""One training step at the cfg5 tensor size (128x160, N=3) with few planes, backward launched eagerly: the target of
`ncu --metrics gpu__time_duration.sum` for the per-kernel breakdown of the backward pass (profiles/r02_backward_kernels.txt)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aarmvsnet_b200 as aa  # noqa: E402
from aarmvsnet_b200.backward import Backward  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H, W, N = 128, 160, 3
sd = torch.load(os.path.join(ROOT, "tests", "golden", "model_dtu_v2.ckpt"), map_location="cpu", weights_only=False)["model"]
m = aa.EMVSNet(32, image_scale=1.0, max_h=H, max_w=W, return_depth=False)
m.load_state_dict(sd, strict=True)
m = m.cuda().train()
s = aa.synth.make_sample(H, W, N, D, seed=3, depth_mode="inverse")
saved = m._engine.forward_saving_state(s["imgs"][0].cuda(), s["proj_matrices"][0], s["depth_values"][0].cuda())
torch.cuda.synchronize()
bwd = Backward(m._engine, torch.device("cuda", 0), use_graph=False)
dcost = torch.randn(D, H * W, device="cuda") * 1e-4
torch.cuda.nvtx.range_push("backward")
bwd.sample(saved, dcost)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done")
