"""Where a cfg5 training step spends its time (128x160, N=3, D=192): the reverse sweep of B samples, host enqueue time against
device time, for B = 1 (the dependent chain of one sample) and B = 4 (the bench's batch).
    python scripts/train_timing.py [D]"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import aarmvsnet_b200 as aa  # noqa: E402
from aarmvsnet_b200.backward import Backward  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 192
H, W, N = 128, 160, 3
sd = torch.load(os.path.join(ROOT, "tests", "golden", "model_dtu_v2.ckpt"), map_location="cpu", weights_only=False)["model"]
m = aa.EMVSNet(32, image_scale=1.0, max_h=H, max_w=W, return_depth=False)
m.load_state_dict(sd, strict=True)
m = m.cuda().train()
eng = m._engine
for B in (1, 2, 4):
    samples = [aa.synth.make_sample(H, W, N, D, seed=3 + b, depth_mode="inverse") for b in range(B)]
    for rep in range(3):
        t0 = time.perf_counter()
        saved = [eng.forward_saving_state(s["imgs"][0].cuda(), s["proj_matrices"][0], s["depth_values"][0].cuda()) for s in samples]
        t_fh = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_f = time.perf_counter() - t0
        bwd = Backward(eng, torch.device("cuda", 0))
        dcost = torch.randn(B, D, H * W, device="cuda") * 1e-4
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bwd.batch(saved, dcost)
        t_host = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
    print(f"B={B}: forward host {t_fh * 1e3:.1f} ms, done {t_f * 1e3:.1f} ms; backward host enqueue {t_host * 1e3:.1f} ms, done {t_all * 1e3:.1f} ms "
          f"({t_all / (B * D) * 1e6:.0f} us per plane and sample)")
