"""Known-answer tests of the tcgen05 / TMEM building blocks (C ABI: aarmvs_tc_selftest)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,group_rows,shift,a_rows", [(16, 8, 0, 128), (48, 8, 0, 128), (48, 10, 0, 180), (48, 10, 11, 180),
                                                       (96, 10, 21, 186), (32, 10, 22, 186)])
def test_umma_unswizzled_shifted_window(K, group_rows, shift, a_rows):
    from aarmvsnet_b200 import capi
    lib = capi.load()
    g = torch.Generator().manual_seed(K * 1000 + shift)
    a = torch.randn(a_rows, K, generator=g).half().cuda()
    b = torch.randn(64, K, generator=g).half().cuda()
    d = torch.zeros(128, 64, device="cuda")
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.check(lib.aarmvs_tc_selftest(ctypes.c_void_p(a.data_ptr()), a_rows, ctypes.c_void_p(b.data_ptr()),
                                      ctypes.c_void_p(d.data_ptr()), K, group_rows, shift, st), "aarmvs_tc_selftest")
    torch.cuda.synchronize()
    m = torch.arange(128)
    rows = shift + (m // 8) * group_rows + m % 8
    want = a.float().cpu()[rows] @ b.float().cpu().t()
    torch.testing.assert_close(d.cpu(), want, rtol=1e-4, atol=1e-3)
