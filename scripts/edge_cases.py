"""Debug: small / odd shapes through the whole forward against the CPU oracle (same winning plane, confidence)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aarmvsnet_b200 as aa
from oracle import aarmvs_oracle as orc
ck = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_dtu_v2.ckpt"), map_location="cpu", weights_only=False)
for (H, W, N, D, pb) in ((16, 16, 2, 1, 8), (16, 24, 3, 3, 8), (20, 12, 2, 9, 4), (8, 8, 3, 5, 2), (28, 36, 9, 17, 16), (64, 40, 16, 4, 3)):
    m = aa.EMVSNet(32, image_scale=1.0, max_h=H, max_w=W, return_depth=True, plane_batch=pb)
    m.load_state_dict(ck["model"], strict=True)
    m = m.cuda().eval()
    s = aa.synth.make_sample(H, W, N, D, seed=H * W + N)
    try:
        out = m(s["imgs"].cuda(), s["proj_matrices"].cuda(), s["depth_values"].cuda())
        torch.cuda.synchronize()
        out2 = m(s["imgs"].cuda(), s["proj_matrices"].cuda(), s["depth_values"].cuda())      # second call: graph replay path
        torch.cuda.synchronize()
        ref = orc.forward_test(ck["model"], s["imgs"], s["proj_matrices"], s["depth_values"], return_index=True)
        same = (out["index"].cpu().long() == ref["index"]).float().mean().item()
        same2 = (out2["index"].cpu().long() == ref["index"]).float().mean().item()
        dconf = (out["photometric_confidence"].cpu() - ref["photometric_confidence"]).abs().max().item()
        print(f"{H}x{W} N={N} D={D} K={pb}: same plane {same * 100:.2f}% / {same2 * 100:.2f}%, max |dconf| {dconf:.2e}")
    except Exception as e:
        print(f"{H}x{W} N={N} D={D} K={pb}: FAILED {type(e).__name__}: {e}")
