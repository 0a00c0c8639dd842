"""Debug: per-tile timeline (clock64) of the loader / MMA / epilogue roles of CTA 0 of ConvLSTM cell 0 at cfg2."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aarmvsnet_b200 as aa
from aarmvsnet_b200 import capi, ops
lib = capi.load()
H, W = 1184, 1600
m = aa.EMVSNet(32, image_scale=1.0, max_h=H, max_w=W, return_depth=True).cuda().eval()
x = torch.randn(32, H, W, device="cuda"); h = torch.randn(16, H, W, device="cuda"); c = torch.randn(16, H, W, device="cuda")
buf = torch.zeros(64 * 12, dtype=torch.int64, device="cuda")
ops.convlstm_cell(m, 0, x, h, c)
lib.aarmvs_debug_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
ops.convlstm_cell(m, 0, x, h, c)
torch.cuda.synchronize()
lib.aarmvs_debug_profile_buffer(None)
t = buf.view(64, 12).cpu()
t0 = t[0, 0].item()
names = ["L.wait0", "L.wait1", "L.stored", "L.arrived", "M.wait0", "M.full", "M.tmem", "M.issued", "E.wait0", "E.full", "E.ld", "E.done"]
print(" tile " + " ".join(f"{n:>9s}" for n in names))
for i in range(24):
    print(f"{i:5d} " + " ".join(f"{(t[i, k].item() - t0):9d}" for k in range(12)))
d = t[8:40]
print("avg per tile: loader fill (wait1->stored) %.0f, loader wait %.0f | mma wait full %.0f, wait tmem %.0f, issue %.0f | epi wait %.0f, ld %.0f, math+store %.0f | tile period %.0f" % (
    (d[:, 2] - d[:, 1]).float().mean(), (d[:, 1] - d[:, 0]).float().mean(), (d[:, 5] - d[:, 4]).float().mean(), (d[:, 6] - d[:, 5]).float().mean(),
    (d[:, 7] - d[:, 6]).float().mean(), (d[:, 9] - d[:, 8]).float().mean(), (d[:, 10] - d[:, 9]).float().mean(), (d[:, 11] - d[:, 10]).float().mean(),
    (d[1:, 7] - d[:-1, 7]).float().mean()))

for flags, name in ((0, "full"), (1, "no loader LDG"), (2, "no loader STS"), (3, "no loader LDG/STS"), (4, "no epilogue global I/O"), (7, "MMA + handshakes only"), (8, "no MMA"), (15, "handshakes only")):
    lib.aarmvs_debug_cell_flags(flags)
    ops.convlstm_cell(m, 0, x, h, c); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ops.convlstm_cell(m, 0, x, h, c)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:26s} {e0.elapsed_time(e1) / 5:.3f} ms  (incl. the NCHW<->NHWC conversions of the seam, constant)")
lib.aarmvs_debug_cell_flags(0)
