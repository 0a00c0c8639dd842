"""Repacks checkpoint tensors (nn.Conv2d / ConvTranspose2d / GroupNorm layouts) into the kernel-native layouts
documented in include/aarmvs.h.  Pure tensor reshuffling; runs once per weight version."""
import torch


def conv_taps(w, cin_pad=None, cout_pad=None):
    """nn.Conv2d weight [cout][cin][k][k] -> [k*k][cin_pad][cout_pad] (zero padded)."""
    cout, cin, k, _ = w.shape
    cin_pad = cin_pad or cin
    cout_pad = cout_pad or cout
    out = w.new_zeros(k * k, cin_pad, cout_pad)
    out[:, :cin, :cout] = w.permute(2, 3, 1, 0).reshape(k * k, cin, cout)
    return out.contiguous()


def pad_vec(b, n):
    out = b.new_zeros(n)
    out[: b.numel()] = b
    return out


def lstm_gates(w, b):
    """ConvLSTMCell conv: output channels [i | f | o | g] blocks of `hidden` (models/module.py:83) ->
    [k*k][cin][hidden*4] with channel order [hidden channel][i,f,o,g], and the bias likewise."""
    c4, cin, k, _ = w.shape
    hidden = c4 // 4
    perm = torch.arange(c4, device=w.device).view(4, hidden).t().reshape(-1)      # new index hc*4+g <- old g*hidden+hc
    return conv_taps(w[perm]), b[perm].contiguous()


def lstm_gates_f16_steps(w, b):
    """The same operand arranged by K STEP, [steps][2 chunks][2 (hi, lo)][4*hidden][8] fp16, as csrc/convlstm_tma.cu issues
    them: per tap one step per 16 input channels; when cin % 16 == 8 (cell 4: 40 channels) the last chunk of taps
    (0,1) (2,3) (4,5) (6,7) (8,-) shares one step each (9 * 2 + 5 = 23 steps instead of 27 zero-padded ones).
    (A handful of whole-tensor ops: a training loop re-packs after every optimizer step.)"""
    taps, bias = lstm_gates(w, b)                    # [9][cin][n] fp32
    k, cin, n = taps.shape
    full = cin // 16

    def split(blk):                                  # [..., 8 channels, n] fp32 -> [..., 2 (hi, lo), n, 8] fp16
        hi = blk.half()
        lo = (blk - hi.float()).half()
        return torch.stack([hi, lo], -3).transpose(-1, -2)

    # regular steps: tap t, 16-channel group s -> chunks (2s, 2s + 1)
    steps = split(taps[:, :16 * full].reshape(k, full, 2, 8, n)).reshape(k * full, 2, 2, n, 8)
    if cin % 16:
        tail = taps.new_zeros(k + 1, 8, n)           # chunk `2 * full` of every tap, plus a zero tap to pair tap 8 with
        tail[:k] = taps[:, 16 * full:16 * full + 8]
        steps = torch.cat([steps, split(tail.reshape((k + 1) // 2, 2, 8, n))], 0)
    return steps.contiguous(), bias


def conv_f16_steps(taps, n_pad=None):
    """[9][cin][n] fp32 tap-major weights (cin a multiple of 16) -> [9 * cin/16 K steps][2 chunks][2 (hi, lo)][n_pad][8] fp16,
    the B operand of csrc/conv_dx_tc.cu (the arrangement of lstm_gates_f16_steps without the gate reordering; output rows
    beyond n are zero)."""
    k, cin, n = taps.shape
    n_pad = n_pad or n
    full = taps.new_zeros(k, cin, n_pad)
    full[:, :, :n] = taps
    blk = full.reshape(k, cin // 16, 2, 8, n_pad)
    hi = blk.half()
    lo = (blk - hi.float()).half()
    return torch.stack([hi, lo], -3).transpose(-1, -2).reshape(k * (cin // 16), 2, 2, n_pad, 8).contiguous()


def taps_f16(taps, n_pad=None):
    """[9][cin][n] fp32 tap-major weights -> tcgen05 B operand [2 planes hi,lo][9][cin_pad/8][n_pad][8] fp16
    (un-swizzled K-major: 16-byte chunks of 8 input channels; cin padded to a multiple of 16, n to n_pad, with zeros)."""
    k, cin, n = taps.shape
    cin_pad = (cin + 15) // 16 * 16
    n_pad = n_pad or n
    full = taps.new_zeros(k, cin_pad, n_pad)
    full[:, :cin, :n] = taps
    chunks = full.view(k, cin_pad // 8, 8, n_pad).permute(0, 1, 3, 2).contiguous()      # [9][kc8][n][8]
    hi = chunks.half()
    lo = (chunks - hi.float()).half()
    return torch.stack([hi, lo]).contiguous()


def stem_taps_f16(w):
    """Omega stem conv [4][32][3][3] as the B operand of pass A's implicit GEMM (csrc/cost_stem_tc.cu):
    [9 taps][4 chunks of 8 input channels][16 rows][8] fp16, rows 0-3 = hi(w[o]), rows 8-11 = lo(w[o]), others zero."""
    cout, cin, k, _ = w.shape
    taps = w.permute(2, 3, 1, 0).reshape(k * k, cin, cout)                # [tap][c][o]
    hi = taps.half()
    lo = (taps - hi.float()).half()
    out = torch.zeros(k * k, cin // 8, 16, 8, dtype=torch.float16, device=w.device)
    out[:, :, 0:cout] = hi.view(k * k, cin // 8, 8, cout).permute(0, 1, 3, 2)
    out[:, :, 8:8 + cout] = lo.view(k * k, cin // 8, 8, cout).permute(0, 1, 3, 2)
    return out.contiguous()


def deconv_taps(w):
    """nn.ConvTranspose2d weight [cin][cout][kh][kw] -> [kh][kw][cin][cout]."""
    return w.permute(2, 3, 0, 1).contiguous()


def deform_taps_f16(w):
    """DeformConv2d.conv weight [32 out][32 in][3][3] -> B operand of csrc/deform_tc.cu:
    [9 taps][2 K steps of 16 input channels][2 chunks][64 rows: 0-31 hi(w[o]), 32-63 lo][8] fp16."""
    cout, cin, k, _ = w.shape
    taps = w.permute(2, 3, 1, 0).reshape(k * k, cin, cout).float()      # [tap][c][o]
    hi = taps.half()
    lo = (taps - hi.float()).half()
    out = torch.zeros(k * k, cin // 16, 2, 2 * cout, 8, dtype=torch.float16, device=w.device)
    out[:, :, :, :cout] = hi.view(k * k, cin // 16, 2, 8, cout).permute(0, 1, 2, 4, 3)
    out[:, :, :, cout:] = lo.view(k * k, cin // 16, 2, 8, cout).permute(0, 1, 2, 4, 3)
    return out.contiguous()


# (ky, kx) of the nine (output parity class, tap) pairs of csrc/deconv_tc.cu, in its issue order
DECONV_PAIRS = ((1, 1), (1, 2), (1, 0), (2, 1), (0, 1), (2, 2), (2, 0), (0, 2), (0, 0))


def deconv_pairs_f16(w):
    """nn.ConvTranspose2d weight [cin=16][cout=16][3][3] -> B operands of the tensor-core deconvolution:
    [9 pairs][2 chunks of 8 input channels][32 rows: 0-15 hi(W[.][co]), 16-31 lo][8] fp16."""
    cin, cout = w.shape[:2]
    out = torch.zeros(len(DECONV_PAIRS), cin // 8, 2 * cout, 8, dtype=torch.float16, device=w.device)
    for p, (ky, kx) in enumerate(DECONV_PAIRS):
        m = w[:, :, ky, kx].float()                          # [ci][co]
        hi = m.half()
        lo = (m - hi.float()).half()
        out[p, :, :cout] = hi.view(cin // 8, 8, cout).permute(0, 2, 1)
        out[p, :, cout:] = lo.view(cin // 8, 8, cout).permute(0, 2, 1)
    return out.contiguous()


# taps (ky * 3 + kx) of the five two-tap MMAs of csrc/head_tc.cu; -1 = zero weights
HEAD_PAIRS = ((0, 1), (2, 3), (4, 5), (6, 7), (8, -1))


def head_pairs_f16(w):
    """conv_0 weight [1][8][3][3] -> [5 pairs][2 taps][16 rows][8] fp16, row 0 = hi(w[tap][c]), row 8 = lo, others zero."""
    taps = w[0].permute(1, 2, 0).reshape(9, 8).float()          # [tap][c]
    hi = taps.half()
    lo = (taps - hi.float()).half()
    out = torch.zeros(len(HEAD_PAIRS), 2, 16, 8, dtype=torch.float16, device=w.device)
    for p, pair in enumerate(HEAD_PAIRS):
        for j, t in enumerate(pair):
            if t >= 0:
                out[p, j, 0] = hi[t]
                out[p, j, 8] = lo[t]
    return out.contiguous()


def omega_tail(sd, prefix="omega.reweight_network"):
    """Flat parameter block in the order of struct OmegaTail (csrc/warp_aggregate.cu)."""
    s = prefix + ".1.stem"
    parts = [
        sd[prefix + ".0.1.weight"], sd[prefix + ".0.1.bias"],
        sd[s + ".0.0.weight"].reshape(16), sd[s + ".0.0.bias"],
        sd[s + ".0.1.weight"], sd[s + ".0.1.bias"],
        sd[s + ".1.weight"].reshape(16), sd[s + ".1.bias"],
        sd[s + ".2.weight"], sd[s + ".2.bias"],
        sd[prefix + ".2.weight"].reshape(4), sd[prefix + ".2.bias"],
    ]
    return torch.cat([p.reshape(-1).float().cpu() for p in parts])
