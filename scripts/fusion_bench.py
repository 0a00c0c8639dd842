"""Measurement of the fusion filter: aarmvs_fuse_view at cfg2 size with 10 source views, CUDA events,
against the numpy oracle (= the reference's algorithm) on the host cores.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from aarmvsnet_b200 import fusion
from oracle import fusion_oracle as fo
from oracle.make_golden_fusion import scene

H, W, S = 1184, 1600, 10
Ks, Es, depths, conf = scene(H, W, S + 1, seed=3, noise=0.003)
d_ref, cf = torch.from_numpy(depths[0]).cuda(), torch.from_numpy(conf).cuda()
d_src = torch.from_numpy(np.stack(depths[1:])).cuda()
run = lambda: fusion.filter_view(d_ref, cf, Ks[0], Es[0], d_src, Ks[1:], Es[1:], 0.35)      # noqa: E731
for _ in range(3):
    out = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
t0 = time.time()
ref = fo.filter_view(depths[0], conf, Ks[0], Es[0], depths[1:3], Ks[1:3], Es[1:3], 0.35)     # 2 of the 10 views
cpu_s = (time.time() - t0) * S / 2
alg = 4 * H * W * (S + 2) + 10 * H * W                       # ref depth + conf + S source maps read, 3 maps written
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
print(json.dumps({"metric": "fusion: reference views/s (1600x1184, 10 source views)", "value": 1e3 / ms, "ms_per_view": ms,
                  "includes": "host-side camera matrices + H2D of 4.8 KB + kernel", "algorithmic_bytes": alg,
                  "achieved_gbs": alg / ms / 1e6, "hbm_peak_gbs": peak, "frac": alg / ms / 1e6 / peak,
                  "cpu_oracle_views_per_s": 1.0 / cpu_s, "cpu_sample": "numpy oracle on 2 of the 10 source views, extrapolated",
                  "geo_mask_mean": float(out["geo_mask"].float().mean())}))
