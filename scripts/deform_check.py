"""Debug / measurement: the deformable conv on the tensor cores (aarmvs_deform_conv_tc) against the fp32 CUDA-core kernel
(aarmvs_deform_conv) at the three FeatNet resolutions of cfg2, with GroupNorm'd input and GN sums; then FeatNet per view."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aarmvsnet_b200 as aa
from aarmvsnet_b200 import capi
from aarmvsnet_b200.engine import _ptr, _src, _src_array

H0, W0 = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1184, 1600)
ck = torch.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "model_dtu_v2.ckpt"), map_location="cpu", weights_only=False)
m = aa.EMVSNet(32, image_scale=1.0, max_h=H0, max_w=W0, return_depth=True)
m.load_state_dict(ck["model"], strict=True)
m = m.cuda().eval()
eng = m._engine
lib = eng.lib
dev = torch.device("cuda")
w = eng._weights(dev)
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for l, s in enumerate((1, 2, 4)):
    H, W = H0 // s, W0 // s
    P = H * W
    x = torch.randn(P, 32, device=dev) * 2 + 0.3
    xs = torch.stack((x.double().view(P, 4, 8).sum((0, 2)), (x.double() ** 2).view(P, 4, 8).sum((0, 2))), 1).contiguous().view(-1)
    src = _src(x, 32, capi.SRC_GN_RELU, sums=xs, gn=w[f"conv{l}.gn"], groups=4, count=8 * P)
    om = torch.empty(P, 32, device=dev)
    capi.check(lib.aarmvs_conv2d(_src_array(src), 1, H, W, 32, 32, 3, 1, _ptr(w[f"pm{l}.w"]), _ptr(w[f"pm{l}.b"]), _ptr(om), 32,
                                 None, st), "pm")
    om[:, :18] += torch.randn(P, 18, device=dev) * 1.5           # spread the offsets (the trained ones are small)
    outs, sums, ms = {}, {}, {}
    for math in ("fma", "tc"):
        out = torch.empty(P, 32, device=dev)
        gs = torch.zeros(8, device=dev, dtype=torch.float64)
        fn_name = "aarmvs_deform_conv_tc" if math == "tc" else "aarmvs_deform_conv"
        wk = w[f"dc{l}.wtc"] if math == "tc" else w[f"dc{l}.w"]

        def run():
            gs.zero_()
            capi.check(getattr(lib, fn_name)(ctypes.byref(src), H, W, _ptr(om), _ptr(wk), _ptr(w[f"dc{l}.b"]), _ptr(out), _ptr(gs), st),
                       fn_name)
        ms[math] = timed(run)
        outs[math], sums[math] = out.clone(), gs.clone()
    d = (outs["tc"] - outs["fma"]).abs()
    print(f"level {l} {H}x{W}: fma {ms['fma']:.3f} ms  tc {ms['tc']:.3f} ms   max|d| {d.max().item():.3e} (ref max {outs['fma'].abs().max().item():.2f})"
          f"  sums rel {((sums['tc'] - sums['fma']).abs() / sums['fma'].abs().clamp_min(1)).max().item():.2e}", flush=True)

img = torch.rand(3, H0, W0, device=dev)
for math in ("fma", "tc"):
    eng.cell_math = math
    t = timed(lambda: m.features(img), reps=3)
    print(f"FeatNet per view ({math} deformable conv): {t:.2f} ms")
