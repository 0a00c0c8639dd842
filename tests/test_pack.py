"""Host-side weight repacking (aarmvsnet_b200/pack.py): the fp16 (hi, lo) tensor-core operands must reproduce the
checkpoint tensors (hi + lo = w to fp16-pair accuracy) in the layouts documented in include/aarmvs.h."""
import torch

from aarmvsnet_b200 import pack


def _close(a, b):
    torch.testing.assert_close(a, b, rtol=0, atol=2e-7 * max(1.0, float(b.abs().max())))


def test_lstm_gate_steps_layout():
    gen = torch.Generator().manual_seed(0)
    for cin, hid in ((48, 16), (32, 16), (40, 8)):
        w = torch.randn(4 * hid, cin, 3, 3, generator=gen)
        b = torch.randn(4 * hid, generator=gen)
        steps, bias = pack.lstm_gates_f16_steps(w, b)
        n = 4 * hid
        assert steps.shape == ((9 * (cin // 16) + (5 if cin % 16 else 0)), 2, 2, n, 8)
        taps, bias_ref = pack.lstm_gates(w, b)                     # [9][cin][n], gate order [hidden][i,f,o,g]
        assert torch.equal(bias, bias_ref)
        full = cin // 16
        rec = (steps[:, :, 0].float() + steps[:, :, 1].float())    # [steps][2 chunks][n][8]
        for t in range(9):
            for s in range(full):
                for c in range(2):
                    _close(rec[t * full + s, c].t(), taps[t, 16 * s + 8 * c:16 * s + 8 * c + 8])
        if cin % 16:                                               # fifth chunk of the tap pairs (0,1) (2,3) (4,5) (6,7) (8,-)
            for p in range(5):
                _close(rec[9 * full + p, 0].t(), taps[2 * p, 16 * full:])
                if 2 * p + 1 < 9:
                    _close(rec[9 * full + p, 1].t(), taps[2 * p + 1, 16 * full:])
                else:
                    assert not rec[9 * full + p, 1].any()
        # gate order: output channel hc * 4 + g of the packed operand is row g * hidden + hc of the checkpoint
        _close(taps[4, :, 1 * 4 + 2], w[2 * hid + 1, :, 1, 1])


def test_stem_deconv_head_operands():
    gen = torch.Generator().manual_seed(1)
    w = torch.randn(4, 32, 3, 3, generator=gen)
    op = pack.stem_taps_f16(w)                                     # [9][4][16][8]: rows 0-3 hi, 8-11 lo
    assert op.shape == (9, 4, 16, 8) and not op[:, :, 4:8].any() and not op[:, :, 12:].any()
    rec = op[:, :, 0:4].float() + op[:, :, 8:12].float()          # [tap][chunk][o][8]
    _close(rec.permute(2, 1, 3, 0).reshape(4, 32, 3, 3), w)
    wd = torch.randn(16, 16, 3, 3, generator=gen)                  # ConvTranspose2d [cin][cout][ky][kx]
    od = pack.deconv_pairs_f16(wd)
    assert od.shape == (9, 2, 32, 8)
    for p, (ky, kx) in enumerate(pack.DECONV_PAIRS):
        rec = od[p, :, :16].float() + od[p, :, 16:].float()       # [chunk][co][8 ci]
        _close(rec.permute(0, 2, 1).reshape(16, 16), wd[:, :, ky, kx])
    wh = torch.randn(1, 8, 3, 3, generator=gen)
    oh = pack.head_pairs_f16(wh)
    assert oh.shape == (5, 2, 16, 8)
    for p, pair in enumerate(pack.HEAD_PAIRS):
        for j, t in enumerate(pair):
            rec = oh[p, j, 0].float() + oh[p, j, 8].float()
            if t < 0:
                assert not oh[p, j].any()
            else:
                _close(rec, wh[0, :, t // 3, t % 3])


def test_deform_taps_layout():
    """B operand of csrc/deform_tc.cu: [9 taps][2 K steps][2 chunks][64 rows: hi(w[o]) | lo(w[o])][8 channels]."""
    gen = torch.Generator().manual_seed(1)
    w = torch.randn(32, 32, 3, 3, generator=gen)
    b = pack.deform_taps_f16(w)
    assert b.shape == (9, 2, 2, 64, 8) and b.dtype == torch.float16
    rec = b[:, :, :, :32].float() + b[:, :, :, 32:].float()          # [tap][step][chunk][o][8]
    for t in range(9):
        ky, kx = divmod(t, 3)
        for s in range(2):
            for c in range(2):
                c0 = 16 * s + 8 * c
                _close(rec[t, s, c], w[:, c0:c0 + 8, ky, kx])
    # hi carries the value to fp16 precision, lo only the remainder
    assert (b[:, :, :, 32:].float().abs() <= b[:, :, :, :32].float().abs() * 2 ** -10 + 1e-7).all()


def test_data_gradient_operand_reproduces_the_transposed_convolution():
    """pack.conv_f16_steps of the spatially flipped W^T (the B operand of csrc/conv_dx_tc.cu): layout [9 * cin/16][2][2][n_pad][8],
    hi + lo = weight, zero rows beyond n, and -- read as a regular 3x3 convolution of the gate gradients, which is what the
    kernel computes -- it is the data gradient of the cell's convolution (autograd)."""
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(5)
    for gates, cin_cell, n_pad in ((64, 48, 48), (64, 32, 32), (32, 40, 48)):
        wt = torch.randn(gates, cin_cell, 3, 3, generator=gen, dtype=torch.float64)            # the cell's conv weight
        wT = wt.permute(2, 3, 0, 1).reshape(9, gates, cin_cell).float()                        # [tap][gate][input]
        steps = pack.conv_f16_steps(wT.flip(0), n_pad)
        assert steps.shape == (9 * gates // 16, 2, 2, n_pad, 8) and steps.dtype == torch.float16
        rec = steps[:, :, 0].float() + steps[:, :, 1].float()                                  # [steps][2 chunks][n_pad][8]
        assert not rec[:, :, cin_cell:].any()
        taps = rec.reshape(9, gates // 16, 2, n_pad, 8).permute(0, 1, 2, 4, 3).reshape(9, gates, n_pad)[:, :, :cin_cell]
        _close(taps, wT.flip(0))
        x = torch.randn(cin_cell, 7, 9, generator=gen, dtype=torch.float64, requires_grad=True)
        y = F.conv2d(x[None], wt, None, padding=1)[0]
        dy = torch.randn(*y.shape, generator=gen, dtype=torch.float64)
        (want,) = torch.autograd.grad(y, x, dy)
        as_conv = taps.double().reshape(3, 3, gates, cin_cell).permute(3, 2, 0, 1)             # [out = input ch][in = gate][ty][tx]
        got = F.conv2d(dy[None], as_conv, None, padding=1)[0]
        torch.testing.assert_close(got, want, rtol=0, atol=1e-5 * float(want.abs().max()))
