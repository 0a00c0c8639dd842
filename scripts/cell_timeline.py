"""Debug: kernel time, ablations and per-tile timeline (clock64) of the TMA producer / MMA / epilogue roles of CTA 0
of one tensor-core ConvLSTM cell at cfg2, called through the C ABI on pre-built operand planes.
usage: cell_timeline.py [cell index 0..4]"""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aarmvsnet_b200 as aa
from aarmvsnet_b200 import capi
from aarmvsnet_b200.engine import _psrc, _psrc_array, _ptr
lib = capi.load()
cell = int(sys.argv[1]) if len(sys.argv) > 1 else 0
H, W = 1184, 1600
scale = (1, 2, 4, 2, 1)[cell]
hid = (16, 16, 16, 16, 8)[cell]
srcs_ch = ((32,), (16,), (16,), (16, 16), (16, 16))[cell]
h, w_ = H // scale, W // scale
P = h * w_
m = aa.EMVSNet(32, image_scale=1.0, max_h=H, max_w=W, return_depth=True).cuda().eval()
wts = m._engine._weights(torch.device("cuda"))
f16 = dict(device="cuda", dtype=torch.float16)
srcs = [torch.randn(2, c // 8, P, 8, **f16) * torch.tensor([1.0, 1e-3], **f16).view(2, 1, 1, 1) for c in srcs_ch]
hp = torch.randn(2, 2, (hid + 7) // 8, P, 8, **f16) * 0.3
cb = torch.randn(lib.aarmvs_cell_state_floats(h, w_, hid), device="cuda")
hf = torch.empty(2, P, hid, device="cuda")
pooled = torch.empty(2, 2, P // 4, 8, **f16) if hid == 16 else None
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
arr = _psrc_array(*[_psrc(s, c) for s, c in zip(srcs, srcs_ch)])
cin = sum(srcs_ch) + hid


def run(d=1):
    capi.check(lib.aarmvs_convlstm_cell_tc(arr, len(srcs), h, w_, cin, hid, _ptr(wts[f"cell{cell}.wtc"]), _ptr(wts[f"cell{cell}.b"]),
                                           _ptr(hp), _ptr(cb), _ptr(hf), _ptr(pooled) if pooled is not None else None, d, None, st), "cell")


buf = torch.zeros(64 * 12, dtype=torch.int64, device="cuda")
run()
lib.aarmvs_debug_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
lib.aarmvs_debug_profile_buffer(None)
t = buf.view(64, 12).cpu()
t0 = t[0, 0].item()
names = ["L.wait0", "L.wait1", "L.issued", "L.issued", "M.wait0", "M.full", "M.tmem", "M.issued", "E.wait0", "E.full", "E.ld", "E.done"]
print(f"cell {cell}: {h}x{w_}, cin {cin}, hidden {hid}")
print(" tile " + " ".join(f"{n:>9s}" for n in names))
for i in range(12):
    print(f"{i:5d} " + " ".join(f"{(t[i, k].item() - t0):9d}" for k in range(12)))
d = t[8:40]
print("avg per tile: TMA issue %.0f, producer wait %.0f | mma wait full %.0f, wait tmem %.0f, issue %.0f | epi wait %.0f, ld %.0f, math+store %.0f | tile period %.0f" % (
    (d[:, 2] - d[:, 1]).float().mean(), (d[:, 1] - d[:, 0]).float().mean(), (d[:, 5] - d[:, 4]).float().mean(), (d[:, 6] - d[:, 5]).float().mean(),
    (d[:, 7] - d[:, 6]).float().mean(), (d[:, 9] - d[:, 8]).float().mean(), (d[:, 10] - d[:, 9]).float().mean(), (d[:, 11] - d[:, 10]).float().mean(),
    (d[1:, 7] - d[:-1, 7]).float().mean()))

for flags, name in ((0, "full"), (1, "no TMA loads"), (4, "no epilogue global I/O"), (5, "MMA + handshakes only"), (8, "no MMA"), (13, "handshakes only"), (16, "no gate transcendentals"), (32, "no max-pool output"), (48, "neither"), (61, "handshakes, no transcendentals, no pool")):
    lib.aarmvs_debug_cell_flags(flags)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(10): run(1 + (i & 1))
    e1.record(); torch.cuda.synchronize()
    print(f"{name:26s} {e0.elapsed_time(e1) / 10:.4f} ms")
lib.aarmvs_debug_cell_flags(0)
