"""Profiling driver: pass A in the production configuration (cfg2, 8 planes per launch, costv stored) and the full-resolution
deformable conv, twice each -- meant to run under `ncu --set full -k regex:cost_stem_tc_kernel|deform_tc_kernel`."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aarmvsnet_b200 as aa
from aarmvsnet_b200 import capi
from aarmvsnet_b200.engine import _ptr, _src
H, W, N, D, K = 1184, 1600, 7, 512, 8
ck = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_dtu_v2.ckpt"), map_location="cpu", weights_only=False)
m = aa.EMVSNet(32, image_scale=1.0, max_h=H, max_w=W, return_depth=True)
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
for _ in range(2):
    eng.cost_stem(_ptr(ws["feat"]), _ptr(ws["feat"], P * 32), P * 32, H, W, S, _ptr(ws["cams"]), w, _ptr(ws["depths"]), 100, None, K,
                  _ptr(ws["t0"]), _ptr(ws["psums"]), st, _ptr(ws["costv"]))
torch.cuda.synchronize()
x = _src(ws["x0"], 32, capi.SRC_GN_RELU, sums=ws["fsums"][6:], gn=w["conv0.gn"], groups=4, count=8 * P)
for _ in range(2):
    capi.check(eng.lib.aarmvs_deform_conv_tc(ctypes.byref(x), H, W, _ptr(ws["om0"]), _ptr(w["dc0.wtc"]), _ptr(w["dc0.b"]), _ptr(ws["dc0"]),
                                             None, st), "deform")
torch.cuda.synchronize()
print("ok")
