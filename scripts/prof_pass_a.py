"""Profiling driver: the plane-sweep stage in the production configuration (cfg2, 8 planes per launch, no per-view cost slab:
passes A and D both gather), twice -- meant to run under
    ncu --set full --clock-control none --import-source on -k regex:"cost_stem_tc_kernel|aggregate_kernel|omega_stats_kernel" -c 8
(profiles/r02_stage_kernels.txt).  `--costv` profiles round 1's variant (pass A stores the cost, pass D streams it)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aarmvsnet_b200 as aa
from aarmvsnet_b200 import capi
from aarmvsnet_b200.engine import _ptr
H, W, N, D, K = 1184, 1600, 7, 512, 8
ck = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_dtu_v2.ckpt"), map_location="cpu", weights_only=False)
m = aa.EMVSNet(32, image_scale=1.0, max_h=H, max_w=W, return_depth=True, keep_costv="--costv" in sys.argv)
m.load_state_dict(ck["model"], strict=True)
m = m.cuda().eval()
eng = m._engine
dev = torch.device("cuda")
w = eng._weights(dev)
ws = eng._workspace(H, W, N, D, dev, False)
s = aa.synth.make_sample(H, W, N, D, seed=0)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for v in range(N):                                    # real FeatNet features (smooth, like the bench), not white noise
    eng.feature_map(s["imgs"][0, v].cuda().contiguous(), ws["feat"][v], ws, w, st)
ws["cams"].copy_(eng.relative_cameras(s["proj_matrices"][0]))
ws["depths"].copy_(s["depth_values"][0])
P, S = H * W, N - 1
costv = _ptr(ws["costv"]) if ws["costv"] is not None else None
ref, src = _ptr(ws["feat"]), _ptr(ws["feat"], P * 32)
for _ in range(2):
    ws["psums"].zero_()
    eng.cost_stem(ref, src, P * 32, H, W, S, _ptr(ws["cams"]), w, _ptr(ws["depths"]), 100, None, K, _ptr(ws["t0"]), _ptr(ws["psums"]), st, costv)
    for stage in (2, 3):
        capi.check(eng.lib.aarmvs_omega_stats(stage, _ptr(ws["t0"]), H, W, S, K, w["tail_host"], _ptr(ws["psums"]), st), "omega_stats")
    capi.check(eng.lib.aarmvs_aggregate(ref, src, P * 32, H, W, S, _ptr(ws["cams"]), w["tail_host"], _ptr(ws["depths"]), 100, None, K,
                                        int(eng.align_corners), _ptr(ws["t0"]), _ptr(ws["psums"]), costv, None, _ptr(ws["costp"]), None, st),
               "aggregate")
torch.cuda.synchronize()
print("ok")
