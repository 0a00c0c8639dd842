"""Debug: tensor-core (TMA-fed) ConvLSTM cells against the fp32-FMA kernel on random inputs, with error statistics."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aarmvsnet_b200 as aa
from aarmvsnet_b200 import ops
torch.manual_seed(0)
m = aa.EMVSNet(32, image_scale=1.0, max_h=64, max_w=64, return_depth=True).cuda().eval()
ck = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_dtu_v2.ckpt"), map_location="cpu", weights_only=False)
m.load_state_dict(ck["model"], strict=True)
for cell, cx, ch in ((0, 32, 16), (1, 16, 16), (3, 32, 16), (4, 32, 8)):
    for H, W in ((32, 40), (36, 52), (20, 100), (128, 160)):
        x = torch.randn(cx, H, W, device="cuda"); h = torch.randn(ch, H, W, device="cuda") * 0.5; c = torch.randn(ch, H, W, device="cuda")
        h1, c1 = ops.convlstm_cell(m, cell, x, h, c, math="tc")
        h2, c2 = ops.convlstm_cell(m, cell, x, h, c, math="fma")
        torch.cuda.synchronize()
        print(f"cell{cell} {H}x{W}: max|dh| {(h1 - h2).abs().max().item():.2e} max|dc| {(c1 - c2).abs().max().item():.2e}  finite {bool(torch.isfinite(h1).all())}")
