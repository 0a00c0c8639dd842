"""Debug: cycles per tcgen05.mma (M=128, K=16) vs N and operand layout, 1 CTA and 148 CTAs."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from aarmvsnet_b200 import capi
lib = capi.load()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
reps = 200
for grid in (1, 148):
    out = torch.zeros(grid, dtype=torch.int64, device="cuda")
    for N in (16, 32, 48, 64, 128, 256):
        for group_rows, shift in ((8, 0), (10, 0), (10, 1)):
            for _ in range(2):
                capi.check(lib.aarmvs_tc_mma_bench(N, group_rows, reps, shift, grid, ctypes.c_void_p(out.data_ptr()), st), "bench")
            torch.cuda.synchronize()
            print(f"grid {grid:3d} N {N:3d} group_rows {group_rows:2d} shift {shift}: {out.float().mean().item() / (reps * 9):7.1f} cycles/MMA")
