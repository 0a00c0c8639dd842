"""Debug: time aarmvs_cost_stem_tc at cfg2 (K = 8 planes) with parts of the kernel switched off."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aarmvsnet_b200 as aa
from aarmvsnet_b200 import capi
from aarmvsnet_b200.engine import _ptr
lib = capi.load()
H, W, N, D, K = 1184, 1600, 7, 512, 8
m = aa.EMVSNet(32, image_scale=1.0, max_h=H, max_w=W, return_depth=True).cuda().eval()
eng = m._engine
w = eng._weights(torch.device("cuda"))
ws = eng._workspace(H, W, N, D, torch.device("cuda"), False)
s = aa.synth.make_sample(H, W, N, D, seed=0)
ws["feat"].normal_()
ws["cams"].copy_(eng.relative_cameras(s["proj_matrices"][0]))
ws["depths"].copy_(s["depth_values"][0])
P, S = H * W, N - 1
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    eng.cost_stem(_ptr(ws["feat"]), _ptr(ws["feat"], P * 32), P * 32, H, W, S, _ptr(ws["cams"]), w, _ptr(ws["depths"]), 100, None, K,
                  _ptr(ws["t0"]), _ptr(ws["psums"]), st)
for flags, name in ((0, "full"), (2, "no recipe math"), (4, "no MMA"), (6, "no recipes, no MMA")):
    lib.aarmvs_debug_flags(flags)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:24s} {e0.elapsed_time(e1) / 3 / K:.3f} ms/plane")
lib.aarmvs_debug_flags(0)

# ---- production configuration: costv stored (mode 1); recipes inline vs precomputed by pass R
def timeit0(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / K
rec_w, rec_base = torch.empty(K, S, P, 4, device="cuda"), torch.empty(K, S, P, device="cuda", dtype=torch.int32)
rec = (_ptr(rec_w), _ptr(rec_base))
def run_a1(recipes=None):
    eng.cost_stem(_ptr(ws["feat"]), _ptr(ws["feat"], P * 32), P * 32, H, W, S, _ptr(ws["cams"]), w, _ptr(ws["depths"]), 100, None, K,
                  _ptr(ws["t0"]), _ptr(ws["psums"]), st, _ptr(ws["costv"]), recipes, sweep=False)
def run_r():
    eng.sweep_recipes(H, W, S, _ptr(ws["cams"]), _ptr(ws["depths"]), 100, None, K, *rec, st)
print(f"pass A, costv stored, recipes inline      {timeit0(run_a1):.3f} ms/plane")
t0_inline = ws["t0"].clone(); cv_inline = ws["costv"][:, :, ::97].clone()
print(f"pass R (sweep_recipes)                    {timeit0(run_r):.3f} ms/plane")
print(f"pass A, costv stored, recipes from pass R {timeit0(lambda: run_a1(rec)):.3f} ms/plane")
print("recipes vs inline: t0 identical", torch.equal(t0_inline, ws["t0"]), " costv identical", torch.equal(cv_inline, ws["costv"][:, :, ::97]))

buf = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
lib.aarmvs_debug_stem_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
run(); torch.cuda.synchronize()
lib.aarmvs_debug_stem_profile_buffer(None)
t = buf.view(64, 8).cpu()[8:56]
t = (t - t[0, 0]).double()
print("group-0 cycles per fill: recipes+bar %.0f | gather+stage %.0f | fence+arrive %.0f | wait MMA %.0f | z rows+bar %.0f | sum+store+stats %.0f | fill period %.0f" % (
    (t[:, 1] - t[:, 0]).mean(), (t[:, 2] - t[:, 1]).mean(), (t[:, 3] - t[:, 2]).mean(), (t[:, 4] - t[:, 3]).mean(), (t[:, 5] - t[:, 4]).mean(),
    (t[:, 6] - t[:, 5]).mean(), (t[1:, 0] - t[:-1, 0]).mean()))

# ---- pass D (aggregate) timing, K planes
cost = torch.empty(K, P, 32, device="cuda")
def run_d():
    capi.check(lib.aarmvs_aggregate(_ptr(ws["feat"]), _ptr(ws["feat"], P * 32), P * 32, H, W, S, _ptr(ws["cams"]), w["tail_host"],
                                    _ptr(ws["depths"]), 100, None, K, 0, _ptr(ws["t0"]), _ptr(ws["psums"]), None, _ptr(cost), None, None, st), "agg")
ws["psums"].fill_(1000.0)
run_d(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): run_d()
e1.record(); torch.cuda.synchronize()
print(f"aggregate                {e0.elapsed_time(e1) / 3 / K:.3f} ms/plane")

# ---- pass D streamed from pass A's stored (warp-ref)^2
cv = torch.randn(K, S, P, 32, device="cuda")
def run_ds():
    capi.check(lib.aarmvs_aggregate(None, None, 0, H, W, S, None, w["tail_host"], None, 100, None, K, 0, _ptr(ws["t0"]), _ptr(ws["psums"]),
                                    _ptr(cv), _ptr(cost), None, None, st), "agg stream")
run_ds(); torch.cuda.synchronize()
e0.record()
for _ in range(3): run_ds()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3 / K
print(f"aggregate (streamed)     {ms:.3f} ms/plane  = {(S * 144 + 128) * P / ms / 1e6:.0f} GB/s")

costp = torch.empty(K, 2, 4, P, 8, device="cuda", dtype=torch.float16)
def run_dp():
    capi.check(lib.aarmvs_aggregate(None, None, 0, H, W, S, None, w["tail_host"], None, 100, None, K, 0, _ptr(ws["t0"]), _ptr(ws["psums"]),
                                    _ptr(cv), None, _ptr(costp), None, st), "agg stream planes")
run_dp(); torch.cuda.synchronize()
e0.record()
for _ in range(3): run_dp()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3 / K
print(f"aggregate (streamed, operand planes out)  {ms:.3f} ms/plane  = {(S * 144 + 128) * P / ms / 1e6:.0f} GB/s")

# ---- split variant: A1 (warp_cost) and A2 (stem conv streaming costv) separately
def timeit(fn, name, reps=3):
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:32s} {e0.elapsed_time(e1) / reps / K:.3f} ms/plane")
ac = 0
def run_a1():
    capi.check(lib.aarmvs_warp_cost(_ptr(ws["feat"]), _ptr(ws["feat"], P * 32), P * 32, H, W, S, _ptr(ws["cams"]), _ptr(ws["depths"]), 100, None, K, ac,
                                    _ptr(cv), st), "a1")
def run_a2():
    capi.check(lib.aarmvs_cost_stem_tc(None, None, 0, H, W, S, None, _ptr(w["stem.wtc"]), w["stem.b_host"], _ptr(ws["depths"]), 100, None, K, ac,
                                       _ptr(ws["t0"]), _ptr(ws["psums"]), _ptr(cv), 2, st), "a2")
timeit(run_a1, "A1 warp_cost (gather + store)")
timeit(run_a2, "A2 stem conv streaming costv")
lib.aarmvs_debug_stem_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
run_a2(); torch.cuda.synchronize()
lib.aarmvs_debug_stem_profile_buffer(None)
t = buf.view(64, 8).cpu()[8:56]
t = (t - t[0, 0]).double()
print("A2 group-0 cycles per fill: recipes+bar %.0f | load+stage %.0f | fence+arrive %.0f | wait MMA %.0f | z rows+bar %.0f | sum+store+stats %.0f | fill period %.0f" % (
    (t[:, 1] - t[:, 0]).mean(), (t[:, 2] - t[:, 1]).mean(), (t[:, 3] - t[:, 2]).mean(), (t[:, 4] - t[:, 3]).mean(), (t[:, 5] - t[:, 4]).mean(),
    (t[:, 6] - t[:, 5]).mean(), (t[1:, 0] - t[:-1, 0]).mean()))
