"""Debug: pass A on the tensor cores (aarmvs_cost_stem_tc) against the fp32-FMA kernel (aarmvs_cost_stem): t0 and GN sums."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aarmvsnet_b200 as aa
from aarmvsnet_b200 import capi
from aarmvsnet_b200.engine import _ptr
H, W, N, D, K = 48, 64, 3, 8, 2
ck = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_dtu_v2.ckpt"), map_location="cpu", weights_only=False)
res = {}
for math in ("tc", "fma"):
    m = aa.EMVSNet(32, image_scale=1.0, max_h=H, max_w=W, return_depth=True, cell_math=math)
    m.load_state_dict(ck["model"], strict=True)
    m = m.cuda().eval()
    eng = m._engine
    dev = torch.device("cuda")
    w = eng._weights(dev)
    s = aa.synth.make_sample(H, W, N, D, seed=0)
    P, S = H * W, N - 1
    torch.manual_seed(1)
    feat = torch.randn(N, P, 32, device=dev)
    cams = eng.relative_cameras(s["proj_matrices"][0]).to(dev)
    depths = s["depth_values"][0].to(dev)
    t0 = torch.zeros(K, S, P, 4, device=dev)
    sums = torch.zeros(K * S * 6 + 8, device=dev, dtype=torch.float64)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    eng.cost_stem(_ptr(feat), _ptr(feat, P * 32), P * 32, H, W, S, _ptr(cams), w, _ptr(depths), 2, None, K, _ptr(t0), _ptr(sums), st)
    torch.cuda.synchronize()
    res[math] = (t0.cpu(), sums.cpu())
a, b = res["tc"][0], res["fma"][0]
d = (a - b).abs().view(K, S, H, W, 4)
print("max |dt0|", d.max().item(), "ref max", b.abs().max().item())
print("per channel", d.amax(dim=(0, 1, 2, 3)))
e = d[0, 0].amax(-1)
print("rows 0..17, cols 0..9 error map (x1e3):")
for y in range(18):
    print(" ".join(f"{e[y, x].item() * 1e3:7.2f}" for x in range(10)))
print("sums tc", res["tc"][1][:4], "fma", res["fma"][1][:4])
for k in range(K):
    for v in range(S):
        print(f"k={k} v={v}: max err {d[k, v].max().item():.3e}   frac bad {(d[k, v].amax(-1) > 1e-3).float().mean().item():.3f}")
kk, vv = 1, 0
e = d[kk, vv].amax(-1)
print(f"error map k={kk} v={vv} rows 0..19 cols 0..17 (x1)")
for y in range(20):
    print(" ".join(f"{e[y, x].item():6.2f}" for x in range(18)))
