"""Backward pass of the training-phase forward (BASELINE config 5; reference: train.py:288-307 `loss.backward()` through
models/drmvsnet.py:272-295): host-side chaining of the fp32 backward kernels behind include/aarmvs.h (csrc/backward*.cu).

Strategy: the forward (Engine.forward_saving_state, tensor-core kernels) keeps the features, the regularised
costs and a snapshot of the recurrent state (h, c of the five ConvLSTM cells) after every depth plane.  The backward walks
the planes in REVERSE (one CUDA graph of ~60 launches, replayed per plane); for plane d it recomputes that plane's
intermediates from the snapshots of planes d-1 and d
(per-view cost, omega's t0 and statistics, aggregated cost, pooled / deconvolved inputs, gate pre-activations), runs the
U-Net's backward against them and carries dL/dh_{d-1}, dL/dc_{d-1} of the five cells to the next iteration.  The warp
backward scatters into per-view feature gradients that are accumulated over all planes; FeatNet's backward then runs once
per view (its forward is recomputed into the engine's scratch buffers first).

Gradient flow follows the reference exactly: the sampling grid is built under no_grad (module.py:15-34), so cameras and
depth hypotheses get no gradient; deformable offsets get gradient through the bilinear weights but not through floor()
(module.py:150); DeformConv2d's backward hooks are no-ops (module.py:128-131).

PyTorch is used for device memory, zero-filling, a handful of adds and the final re-layout of the weight gradients.
"""
import ctypes

import torch

from . import capi, pack
from .engine import _HID, _SCALE, _ptr, _src

_CELL_IN = (32, 16, 16, 32, 32)           # input channels of the five cells (models/drmvsnet.py:241)
_FS_OFF = {"init0": 0, "init1": 2, "conv0": 6, "conv1": 14, "conv2": 22, "dc0": 30, "dc1": 38, "dc2": 46,
           "head0": 54, "head1": 58, "head2": 60}          # GroupNorm sums inside Engine._feat_scratch()["fsums"]


def _seg(t, channels, accumulate=0):
    s = capi.Seg()
    s.ptr = t.data_ptr()
    s.channels, s.accumulate = channels, accumulate
    return s


def _segs(*items):
    return (capi.Seg * len(items))(*items)


def _taps_inverse(dw, cout, cin, k):
    """[k*k][cin_pad][cout_pad] gradient -> nn.Conv2d layout [cout][cin][k][k]."""
    return dw[:, :cin, :cout].reshape(k, k, cin, cout).permute(3, 2, 0, 1).contiguous()


class Backward:
    """Gradient accumulation for one optimisation step of one EMVSNet (all samples of the batch, all planes)."""

    def __init__(self, engine, device, use_graph=True, dx_math="tc"):
        """dx_math: the cells' data gradients on the tensor cores ("tc", csrc/conv_dx_tc.cu) or as fp32 FMA code ("fma", the
        register-tiled conv_dx3_kernel: tests and A/B measurements)."""
        self.eng = engine
        self.use_graph = use_graph
        self.dx_math = dx_math
        self.lib = engine.lib
        self.dev = device
        self.sd = {k: v.detach().to(device=device, dtype=torch.float32) for k, v in engine.module.state_dict().items()}
        self.w = engine._weights(device)
        self._pack()
        # slot -> workspace and stream, one per concurrently running sample (see batch()).  They live on the engine and survive
        # the optimisation step (this object does not: the weights it packs change every step), so that the caching allocator's
        # per-stream pools are reused; only the captured graphs are per step (they bake the packed weights' addresses).
        if not hasattr(engine, "_bwd_state"):
            engine._bwd_state = {}                      # per device: (workspaces by slot, streams)
        self._wss, self._streams = engine._bwd_state.setdefault(str(device), ({}, []))
        for ws in self._wss.values():
            ws["graph"] = None
        self.chan = torch.zeros(128, device=device, dtype=torch.float64)      # per-channel sums of one GroupNorm backward

    # ------------------------------------------------------------------------------------------ weights and accumulators
    def _pack(self):
        sd, dev = self.sd, self.dev
        f32 = dict(device=dev, dtype=torch.float32)
        bw, g = {}, {}
        cr = "cost_regularization"
        for i in range(5):
            wt = pack.conv_taps(sd[f"{cr}.cell_list.{i}.conv.weight"])           # [9][cin][4Ch], PyTorch gate order i|f|o|g
            bw[f"cell{i}.w"] = wt
            bw[f"cell{i}.wT"] = wt.transpose(1, 2).contiguous()                   # [9][4Ch][cin]: dX = dG (*) W^T
            # the same as the B operand of the tensor-core data gradient: taps flipped (its taps index the input offset
            # -1..1, aarmvs_conv_nhwc's `transposed` ones the output offset), output rows padded to 48 / 32
            bw[f"cell{i}.dxtc"] = pack.conv_f16_steps(bw[f"cell{i}.wT"].flip(0), 32 if wt.shape[1] == 32 else 48)
            bw[f"cell{i}.b"] = sd[f"{cr}.cell_list.{i}.conv.bias"].contiguous()
            g[f"cell{i}.w"] = torch.zeros_like(wt)
            g[f"cell{i}.b"] = torch.zeros(wt.shape[2], **f32)
        for i in range(2):
            wt = sd[f"{cr}.deconv_{i}.conv.weight"].permute(2, 3, 0, 1).reshape(9, 16, 16).contiguous()     # [tap][ci][co]
            bw[f"deconv{i}.w"] = wt
            bw[f"deconv{i}.wT"] = wt.transpose(1, 2).contiguous()
            bw[f"deconv{i}.b"] = sd[f"{cr}.deconv_{i}.conv.bias"].contiguous()
            bw[f"deconv{i}.gn"] = (sd[f"{cr}.deconv_{i}.gn.weight"].contiguous(), sd[f"{cr}.deconv_{i}.gn.bias"].contiguous())
            g[f"deconv{i}.wT"] = torch.zeros(9, 16, 16, **f32)                   # [tap][co][ci]
            g[f"deconv{i}.b"] = torch.zeros(16, **f32)
            g[f"deconv{i}.gamma"] = torch.zeros(16, **f32)
            g[f"deconv{i}.beta"] = torch.zeros(16, **f32)
        bw["head.wT"] = pack.conv_taps(sd[cr + ".conv_0.weight"]).transpose(1, 2).contiguous()             # [9][1][8]
        g["head.w"] = torch.zeros(9, 8, 1, **f32)
        g["head.b"] = torch.zeros(1, **f32)
        om = "omega.reweight_network"
        bw["stem.wT"] = pack.conv_taps(sd[om + ".0.0.weight"]).transpose(1, 2).contiguous()                # [9][4][32]
        g["stem.w"] = torch.zeros(9, 32, 4, **f32)
        g["stem.b"] = torch.zeros(4, **f32)
        g["tail"] = torch.zeros(self.lib.aarmvs_omega_tail_floats(), device=dev, dtype=torch.float64)
        # FeatNet: the engine's [tap][cin][cout] packs + their transposes
        for name, cin, cout, k in (("init0", 4, 8, 3), ("init1", 8, 16, 3), ("conv0", 16, 32, 3), ("conv1", 32, 32, 3),
                                   ("conv2", 32, 32, 3), ("head0", 32, 16, 1), ("head1", 32, 8, 1), ("head2", 32, 8, 1),
                                   ("pm0", 32, 32, 3), ("pm1", 32, 32, 3), ("pm2", 32, 32, 3)):
            bw[name + ".wT"] = self.w[name + ".w"].transpose(1, 2).contiguous()
            g[name + ".w"] = torch.zeros(k * k, cin, cout, **f32)
            g[name + ".b"] = torch.zeros(cout, **f32)
            if not name.startswith("pm"):
                g[name + ".gamma"] = torch.zeros(cout, **f32)
                g[name + ".beta"] = torch.zeros(cout, **f32)
        for l in range(3):
            w288 = self.w[f"dc{l}.w"].reshape(1, 288, 32)                        # [(tap, c)][o]: the contraction as a 1x1 conv
            bw[f"dc{l}.w288T"] = w288.transpose(1, 2).contiguous()                # [1][32][288]
            g[f"dc{l}.w"] = torch.zeros(1, 288, 32, **f32)
            g[f"dc{l}.b"] = torch.zeros(32, **f32)
            g[f"dc{l}.gamma"] = torch.zeros(32, **f32)
            g[f"dc{l}.beta"] = torch.zeros(32, **f32)
        self.bw, self.g = bw, g

    def _workspace(self, H, W, N, slot=0):
        key = (H, W, N)
        ws = self._wss.get(slot)
        if ws is not None and ws["key"] == key:
            return ws
        dev = self.dev
        f32 = dict(device=dev, dtype=torch.float32)
        P, S = H * W, N - 1
        Pi = [P // (s * s) for s in _SCALE]
        b = {"key": key}
        b["costv"] = torch.empty(S, P, 32, **f32)
        b["t0"] = torch.empty(S, P, 4, **f32)
        # every fp64 accumulator a plane starts from zero, in one block (one fill per plane): omega's forward / backward sums,
        # the two GroupNorms' sums and the per-channel sums of their backward passes
        z = torch.zeros(2 * S * 6 + 2 * 4 + 2 * 128 + 4, device=dev, dtype=torch.float64)
        b["zeros"] = z
        b["stats"], b["bstats"] = z[:S * 6], z[S * 6:2 * S * 6]
        o = 2 * S * 6
        b["gsum0"], b["gsum1"], b["chan0"], b["chan1"] = z[o:o + 4], z[o + 4:o + 8], z[o + 8:o + 136], z[o + 136:o + 264]
        b["gmax"] = z[o + 264:o + 268].view(torch.int32)       # max |dgates| of the five cells (bit patterns, atomicMax)
        b["cost"] = torch.empty(P, 32, **f32)
        b["rew"] = torch.empty(S, P, **f32)
        b["pool0"] = torch.empty(P // 4, 16, **f32)
        b["pool1"] = torch.empty(P // 16, 16, **f32)
        for i, pp in ((0, P // 4), (1, P)):
            b[f"up{i}raw"] = torch.empty(pp, 16, **f32)
            b[f"up{i}"] = torch.empty(pp, 16, **f32)
            b[f"dup{i}"] = torch.empty(pp, 16, **f32)
        b["dpool0"] = torch.empty(P // 4, 16, **f32)
        b["dpool1"] = torch.empty(P // 16, 16, **f32)
        b["gates_pp"] = torch.zeros(5, device=dev, dtype=torch.int64)     # bases of the sample's five gate histories (forward)
        b["dgates"] = [torch.empty(Pi[i], 4 * _HID[i], **f32) for i in range(5)]
        # the same as scaled fp16 (hi, lo) operand planes for the tensor-core data gradient, and the scale's inverse
        b["dgplanes"] = [torch.empty(2, 4 * _HID[i] // 8, Pi[i], 8, device=dev, dtype=torch.float16) for i in range(5)]
        b["ginv"] = torch.ones(5, **f32)
        # ... and the cells' inputs [x | h_{d-1}] as operand planes for the tensor-core weight gradient
        cxs = (48, 32, 32, 48, 40)
        b["xplanes"] = [torch.empty(2, cxs[i] // 8, Pi[i], 8, device=dev, dtype=torch.float16) for i in range(5)]
        b["dh"] = [torch.zeros(Pi[i], _HID[i], **f32) for i in range(5)]
        b["dc"] = [torch.zeros(Pi[i], _HID[i], **f32) for i in range(5)]
        b["dcost"] = torch.empty(P, 32, **f32)
        b["dcostv"] = torch.empty(S, P, 32, **f32)
        b["drew"] = torch.empty(S, P, **f32)
        b["dt0"] = torch.empty(S, P, 4, **f32)
        b["dfeat"] = torch.zeros(N, P, 32, **f32)
        # staging of one plane: the state records after planes d-1 and d (Engine.state_layout), dL/dcost_reg[d], the plane index
        lay, rec = self.eng.state_layout(H, W)
        b["rec2"] = torch.zeros(2, rec, **f32)
        for r, (hn, cn) in enumerate((("hp", "cp"), ("h", "c"))):
            b[hn] = [b["rec2"][r][lay[i][0]:lay[i][0] + lay[i][2]].view(Pi[i], _HID[i]) for i in range(5)]
            b[cn] = [b["rec2"][r][lay[i][1]:lay[i][1] + lay[i][2]].view(Pi[i], _HID[i]) for i in range(5)]
        b["dcr"] = torch.zeros(P, **f32)
        b["plane"] = torch.zeros(1, device=dev, dtype=torch.int32)
        b["feat"] = self.eng.feature_buffer(N, H, W, dev)
        b["cams"] = torch.empty(S, 12, **f32)
        b["depths"] = torch.empty(0, **f32)
        b["graph"] = None
        b["fchan"] = torch.zeros(128, device=dev, dtype=torch.float64)      # GroupNorm-backward scratch of FeatNet's backward
        b["side"] = torch.cuda.Stream(dev)                                  # second branch of the plane (see _plane)
        b["stage_ready"] = torch.cuda.Event()
        self._wss[slot] = b
        return b

    # ------------------------------------------------------------------------------------------ kernel wrappers
    def _st(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)

    def conv(self, ins, hin, win, outs, hout, wout, k, stride, transposed, weight, bias=None):
        cin = sum(s.channels for s in ins)
        cout = sum(s.channels for s in outs)
        capi.check(self.lib.aarmvs_conv_nhwc(_segs(*ins), len(ins), hin, win, _segs(*outs), len(outs), hout, wout, cin, cout, k, stride,
                                             int(transposed), _ptr(weight), _ptr(bias) if bias is not None else None, self._st()),
                   "aarmvs_conv_nhwc")

    def wgrad(self, xs, hx, wx, dy, hy, wy, cy, k, stride, dw, db=None):
        cx = sum(s.channels for s in xs)
        capi.check(self.lib.aarmvs_conv_bwd_weight(_segs(*xs), len(xs), hx, wx, _ptr(dy), hy, wy, cx, cy, k, stride, _ptr(dw),
                                                   _ptr(db) if db is not None else None, self._st()), "aarmvs_conv_bwd_weight")

    def gn_apply(self, x, P, C, groups, sums, gn, out):
        capi.check(self.lib.aarmvs_gn_relu_apply(_ptr(x), P, C, groups, _ptr(sums), _ptr(gn[0]), _ptr(gn[1]), 1, _ptr(out), self._st()),
                   "aarmvs_gn_relu_apply")

    def gn_bwd(self, x, dy, P, C, groups, sums, gn, dx, dgamma, dbeta, dbias=None, chan=None, zero=False):
        """chan: per-channel fp64 scratch, zeroed by the caller unless `zero` (default: this object's, zeroed here)."""
        if chan is None:
            chan, zero = self.chan, True
        if zero:
            chan.zero_()
        for stage in (1, 2):
            capi.check(self.lib.aarmvs_gn_relu_bwd(stage, _ptr(x), _ptr(dy), P, C, groups, _ptr(sums), _ptr(gn[0]), _ptr(gn[1]), 1,
                                                   _ptr(chan), _ptr(dx), _ptr(dgamma), _ptr(dbeta),
                                                   _ptr(dbias) if dbias is not None else None, self._st()), "aarmvs_gn_relu_bwd")

    # ------------------------------------------------------------------------------------------ one sample
    def batch(self, saved_list, dcost):
        """All samples of a step: dcost [B, D, P].  The samples' reverse sweeps are independent chains of short kernels on small
        maps (latency-bound one at a time), so each runs on its own stream with its own workspace and graph, interleaved plane by
        plane; the gradient accumulators are shared (atomics).  FeatNet's backward follows on the calling stream."""
        B = len(saved_list)
        if B == 1:
            return self.sample(saved_list[0], dcost[0])
        main = torch.cuda.current_stream(self.dev)
        while len(self._streams) < B:
            self._streams.append(torch.cuda.Stream(self.dev))
        streams = self._streams[:B]
        sweeps = []
        for i, (sv, st) in enumerate(zip(saved_list, streams)):
            st.wait_stream(main)
            with torch.cuda.stream(st):
                sweeps.append(self._sweep(sv, dcost[i], slot=i))
        alive = True
        while alive:                              # round-robin: one plane of every sample per turn
            alive = False
            for sweep, st in zip(sweeps, streams):
                with torch.cuda.stream(st):
                    if next(sweep, "done") != "done":
                        alive = True
        # FeatNet's backward (once per view; its forward is recomputed first) follows each sample's sweep on the sample's stream,
        # with the sample's own FeatNet scratch: four chains of short kernels side by side instead of twelve views in a row
        for i, (sv, st) in enumerate(zip(saved_list, streams)):
            with torch.cuda.stream(st):
                b = self._wss[i]
                for v in range(sv["N"]):
                    self.featnet(sv["imgs"][v], b["dfeat"][v], sv["H"], sv["W"], slot=i)
        for st in streams:
            main.wait_stream(st)

    def sample(self, saved, dcost_volume):
        """One sample, on the calling stream: reverse sweep, then FeatNet's backward per view."""
        for _ in self._sweep(saved, dcost_volume, slot=0):
            pass
        b = self._wss[0]
        for v in range(saved["N"]):
            self.featnet(saved["imgs"][v], b["dfeat"][v], saved["H"], saved["W"])

    def _sweep(self, saved, dcost_volume, slot):
        """Generator over the planes of one sample's reverse sweep (yields after enqueueing each plane).
        saved: Engine.forward_saving_state's dict; dcost_volume [D, P] = dL / d(regularised cost).  Accumulates into self.g.

        The body of one plane works on FIXED buffers: the two state records it needs (after planes d-1 and d) and its row of
        dcost_volume are staged into them, the plane index lives in device memory -- so the ~60 launches of a plane are captured
        once as a CUDA graph and replayed for every other plane (three small staging ops + one graph launch per plane)."""
        eng = self.eng
        H, W, N, D = saved["H"], saved["W"], saved["N"], saved["D"]
        b = self._workspace(H, W, N, slot)
        for t in b["dh"] + b["dc"]:
            t.zero_()
        b["dfeat"].zero_()
        dcost_volume = dcost_volume.contiguous()
        state = saved["state"]
        # per-sample constants go to fixed buffers as well, so that ONE captured graph serves every sample of the step
        b["feat"].copy_(saved["feat"])
        b["cams"].copy_(saved["cams"])
        if b["depths"].numel() != D:
            b["depths"], b["graph"] = torch.empty(D, device=self.dev, dtype=torch.float32), None
        b["depths"].copy_(saved["depths"])
        for i, gh in enumerate(saved["gates"]):
            b["gates_pp"][i].fill_(gh.data_ptr())
        for d in range(D - 1, -1, -1):
            b["rec2"].copy_(state[d:d + 2])                   # [0] = state entering plane d, [1] = state after it
            b["dcr"].copy_(dcost_volume[d])
            b["plane"].fill_(d)
            if b["graph"] is not None:
                b["graph"].replay()
                yield d
                continue
            self._plane(H, W, N, b)                           # eager: also sets the kernels' launch attributes
            if self.use_graph:
                # capture on a side stream with the raw API: torch.cuda.graph()'s context manager would also run gc.collect(),
                # empty the caching allocator and synchronise the device on every optimisation step (the body allocates nothing)
                cur = torch.cuda.current_stream(self.dev)
                side = torch.cuda.Stream(self.dev)
                side.wait_stream(cur)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.stream(side):
                    graph.capture_begin()
                    self._plane(H, W, N, b)
                    graph.capture_end()
                cur.wait_stream(side)
                b["graph"] = graph
            yield d

    def _plane(self, H, W, N, b):
        """Backward of the plane whose index is in b["plane"], whose states are in b["rec2"] and whose dL/dcost_reg is b["dcr"]
        (features, cameras and depth hypotheses of the sample in b["feat"], b["cams"], b["depths"])."""
        lib, eng, w, bw, g = self.lib, self.eng, self.w, self.bw, self.g
        P, S = H * W, N - 1
        H2, W2, H4, W4 = H // 2, W // 2, H // 4, W // 4
        hw = [(H // s, W // s) for s in _SCALE]
        st = self._st()
        feat = b["feat"]
        ref, src = _ptr(feat), _ptr(feat, P * 32)
        ac = int(eng.align_corners)
        dptr = _ptr(b["plane"])
        depths, cams = _ptr(b["depths"]), _ptr(b["cams"])
        hp, cp, h, c = b["hp"], b["cp"], b["h"], b["c"]
        # Two branches (captured as a fork / join of the plane's graph): the MAIN stream carries the dependent chain of data
        # gradients, the SIDE stream everything nothing downstream in the plane waits for -- the weight gradients (a third of the
        # kernel time) and the recompute of the warp / omega / aggregation stage, which only cell 0's weight gradient and the
        # aggregation backward at the very end need.
        main, side = torch.cuda.current_stream(self.dev), b["side"]

        def on_side(fn):
            side.wait_stream(main)
            with torch.cuda.stream(side):
                fn()

        # ---------------- recompute the plane's forward intermediates
        b["zeros"].zero_()

        def stage_recompute():
            sst = self._st()
            eng.cost_stem(ref, src, P * 32, H, W, S, cams, w, depths, 0, dptr, 1, _ptr(b["t0"]), _ptr(b["stats"]), sst, _ptr(b["costv"]))
            for stage in (2, 3):
                capi.check(lib.aarmvs_omega_stats(stage, _ptr(b["t0"]), H, W, S, 1, w["tail_host"], _ptr(b["stats"]), sst), "aarmvs_omega_stats")
            capi.check(lib.aarmvs_aggregate(ref, src, P * 32, H, W, S, cams, w["tail_host"], depths, 0, dptr, 1, ac, _ptr(b["t0"]),
                                            _ptr(b["stats"]), _ptr(b["costv"]), _ptr(b["cost"]), None, _ptr(b["rew"]), sst), "aarmvs_aggregate")
            b["stage_ready"].record(side)
            capi.check(lib.aarmvs_maxpool2(_ptr(h[0]), H, W, 16, _ptr(b["pool0"]), sst), "aarmvs_maxpool2")
            capi.check(lib.aarmvs_maxpool2(_ptr(h[1]), H2, W2, 16, _ptr(b["pool1"]), sst), "aarmvs_maxpool2")

        on_side(stage_recompute)
        for i, (hin, win, src_h) in enumerate(((H4, W4, h[2]), (H2, W2, h[3]))):
            self.conv([_seg(src_h, 16)], hin, win, [_seg(b[f"up{i}raw"], 16)], 2 * hin, 2 * win, 3, 2, True, bw[f"deconv{i}.w"],
                      bw[f"deconv{i}.b"])
            capi.check(lib.aarmvs_gn_sums(_ptr(b[f"up{i}raw"]), 4 * hin * win, 16, 2, _ptr(b[f"gsum{i}"]), st), "aarmvs_gn_sums")
            self.gn_apply(b[f"up{i}raw"], 4 * hin * win, 16, 2, b[f"gsum{i}"], bw[f"deconv{i}.gn"], b[f"up{i}"])
        xin = [[_seg(b["cost"], 32)], [_seg(b["pool0"], 16)], [_seg(b["pool1"], 16)], [_seg(b["up0"], 16), _seg(h[1], 16)],
               [_seg(b["up1"], 16), _seg(h[0], 16)]]
        # (the cells' gates are not recomputed: the training forward stored them activated, saved["gates"])

        # ---------------- backward of the plane.  b["dh"][i] enters holding dL/dh_i^d from plane d+1 (cell i's h_prev input)
        dh, dc = b["dh"], b["dc"]
        dcr = b["dcr"]                                               # [P] = [P][1]
        self.conv([_seg(dcr, 1)], H, W, [_seg(dh[4], 8, 1)], H, W, 3, 1, True, bw["head.wT"])
        on_side(lambda: self.wgrad([_seg(h[4], 8)], H, W, dcr, H, W, 1, 3, 1, g["head.w"], g["head.b"]))

        def cell_bwd(i, outs):
            # at plane 0 the staged previous state is the zero record: c_prev = h_prev = 0 (module.py:94-96)
            tc = self.dx_math == "tc"
            xs = xin[i] + [_seg(hp[i], _HID[i])]
            ph, pw = hw[i]
            capi.check(lib.aarmvs_lstm_gates_bwd_act(_ptr(b["gates_pp"], i), 0, dptr, _ptr(cp[i]), _ptr(c[i]), _ptr(dh[i]), _ptr(dc[i]),
                                                     _ptr(b["dgates"][i]), ph * pw, _HID[i], _ptr(b["gmax"], i) if tc else None,
                                                     _ptr(g[f"cell{i}.b"]) if tc else None, st), "aarmvs_lstm_gates_bwd_act")
            # dL/dh_i^{d-1} replaces dh[i]: every contribution to dL/dh_i^d precedes this point of the plane, the ones of plane
            # d-1 (skip connections, pooling, deconvolutions, the head) accumulate onto it in the next replay
            outs = outs + [_seg(dh[i], _HID[i])]
            if not tc:   # fp32 FMA kernels (tests, A/B measurements)
                on_side(lambda: self.wgrad(xs, ph, pw, b["dgates"][i], ph, pw, 4 * _HID[i], 3, 1, g[f"cell{i}.w"], g[f"cell{i}.b"]))
                self.conv([_seg(b["dgates"][i], 4 * _HID[i])], ph, pw, outs, ph, pw, 3, 1, True, bw[f"cell{i}.wT"])
                return
            # tensor cores: the gate gradients as scaled fp16 (hi, lo) operand planes -- the A operand of the data gradient (the
            # forward cell's implicit GEMM with the transposed weights) and, read as an MN-major operand, the B operand of the
            # weight gradient (K = pixels; its A operand: the cell's inputs [x | h_{d-1}] as planes).  The bias gradient came
            # from the gate kernel.
            capi.check(lib.aarmvs_split_planes_scaled(_ptr(b["dgates"][i]), ph, pw, 4 * _HID[i], _ptr(b["gmax"], i), _ptr(b["dgplanes"][i]),
                                                      _ptr(b["ginv"], i), st), "aarmvs_split_planes_scaled")

            def dw_tc():
                sst = self._st()
                capi.check(lib.aarmvs_split_planes_cat(_segs(*xs), len(xs), ph, pw, _ptr(b["xplanes"][i]), sst), "aarmvs_split_planes_cat")
                capi.check(lib.aarmvs_conv_dw_tc(_ptr(b["xplanes"][i]), _ptr(b["dgplanes"][i]), ph, pw, sum(x.channels for x in xs),
                                                 4 * _HID[i], _ptr(b["ginv"], i), _ptr(g[f"cell{i}.w"]), sst), "aarmvs_conv_dw_tc")

            on_side(dw_tc)
            capi.check(lib.aarmvs_conv_dx_tc(_ptr(b["dgplanes"][i]), ph, pw, 4 * _HID[i], _segs(*outs), len(outs),
                                             sum(o.channels for o in outs), _ptr(bw[f"cell{i}.dxtc"]), _ptr(b["ginv"], i), st),
                       "aarmvs_conv_dx_tc")

        def up_bwd(i, hin, win, src_h, dh_target):
            pp = 4 * hin * win
            self.gn_bwd(b[f"up{i}raw"], b[f"dup{i}"], pp, 16, 2, b[f"gsum{i}"], bw[f"deconv{i}.gn"], b[f"dup{i}"], g[f"deconv{i}.gamma"],
                        g[f"deconv{i}.beta"], g[f"deconv{i}.b"], chan=b[f"chan{i}"])
            on_side(lambda: self.wgrad([_seg(b[f"dup{i}"], 16)], 2 * hin, 2 * win, src_h, hin, win, 16, 3, 2, g[f"deconv{i}.wT"]))
            self.conv([_seg(b[f"dup{i}"], 16)], 2 * hin, 2 * win, [_seg(dh_target, 16, 1)], hin, win, 3, 2, False, bw[f"deconv{i}.wT"])

        cell_bwd(4, [_seg(b["dup1"], 16), _seg(dh[0], 16, 1)])
        up_bwd(1, H2, W2, h[3], dh[3])
        cell_bwd(3, [_seg(b["dup0"], 16), _seg(dh[1], 16, 1)])
        up_bwd(0, H4, W4, h[2], dh[2])
        cell_bwd(2, [_seg(b["dpool1"], 16)])
        capi.check(lib.aarmvs_maxpool2_bwd(_ptr(h[1]), _ptr(b["dpool1"]), H2, W2, 16, _ptr(dh[1]), st), "aarmvs_maxpool2_bwd")
        cell_bwd(1, [_seg(b["dpool0"], 16)])
        capi.check(lib.aarmvs_maxpool2_bwd(_ptr(h[0]), _ptr(b["dpool0"]), H, W, 16, _ptr(dh[0]), st), "aarmvs_maxpool2_bwd")
        cell_bwd(0, [_seg(b["dcost"], 32)])

        # ---------------- aggregation, omega, warp
        main.wait_event(b["stage_ready"])                            # costv, t0, stats, rew of the recompute branch
        capi.check(lib.aarmvs_aggregate_bwd(_ptr(b["dcost"]), _ptr(b["costv"]), _ptr(b["rew"]), S, P, _ptr(b["dcostv"]), _ptr(b["drew"]), st),
                   "aarmvs_aggregate_bwd")
        for stage in (3, 2, 1, 0):
            capi.check(lib.aarmvs_omega_bwd(stage, _ptr(b["t0"]), _ptr(b["drew"]), H, W, S, w["tail_host"], _ptr(b["stats"]),
                                            _ptr(b["bstats"]), _ptr(g["tail"]), _ptr(b["dt0"]), st), "aarmvs_omega_bwd")

        def stem_wgrads():
            for v in range(S):
                self.wgrad([_seg(b["costv"][v], 32)], H, W, b["dt0"][v], H, W, 4, 3, 1, g["stem.w"], g["stem.b"])

        on_side(stem_wgrads)
        for v in range(S):
            self.conv([_seg(b["dt0"][v], 4)], H, W, [_seg(b["dcostv"][v], 32, 1)], H, W, 3, 1, True, bw["stem.wT"])
        capi.check(lib.aarmvs_warp_bwd(ref, src, P * 32, H, W, S, cams, depths, 0, dptr, ac, _ptr(b["dcostv"]), _ptr(b["dfeat"]),
                                       _ptr(b["dfeat"], P * 32), st), "aarmvs_warp_bwd")
        # (carry: dh now holds dL/dh_{d-1}, the next (lower) plane's starting point; dc was updated in place)
        main.wait_stream(side)                                       # join: the plane's buffers are free for the next plane

    # ------------------------------------------------------------------------------------------ FeatNet of one view
    def featnet(self, img, dfeat, H, W, slot=0):
        """img [3,H,W]; dfeat [P,32] = dL/d(feature map of this view).  slot: which FeatNet scratch / reduction scratch to use
        (concurrent samples)."""
        lib, eng, w, bw, g = self.lib, self.eng, self.w, self.bw, self.g
        st = self._st()
        dev = self.dev
        f32 = dict(device=dev, dtype=torch.float32)
        P = H * W
        fs = eng._feat_scratch(H, W, dev, slot)
        chan = self._wss[slot]["fchan"] if slot in self._wss else self.chan
        scratch = torch.empty(P, 32, **f32)
        eng.feature_map(img, scratch, fs, w, st)               # recompute: raw activations + GroupNorm sums in fs
        sums = fs["fsums"]

        def sl(name):
            return sums[_FS_OFF[name]:]

        lv = [(H, W), (H // 2, W // 2), (H // 4, W // 4)]
        # post-GroupNorm+ReLU activations the convolutions read (materialised: the weight gradients need them)
        a1p, a2p = torch.empty(P, 8, **f32), torch.empty(P, 16, **f32)
        self.gn_apply(fs["a1"], P, 8, 1, sl("init0"), w["init0.gn"], a1p)
        self.gn_apply(fs["a2"], P, 16, 2, sl("init1"), w["init1.gn"], a2p)
        xp, dxp = [], []
        for l, (hl, wl) in enumerate(lv):
            t = torch.empty(hl * wl, 32, **f32)
            self.gn_apply(fs[f"x{l}"], hl * wl, 32, 4, sl(f"conv{l}"), w[f"conv{l}.gn"], t)
            xp.append(t)
            dxp.append(torch.zeros(hl * wl, 32, **f32))
        # IntraViewAA tail
        dm = [torch.empty(P, 16, **f32), torch.zeros(P // 4, 8, **f32), torch.zeros(P // 16, 8, **f32)]
        capi.check(lib.aarmvs_feat_assemble_bwd(_ptr(dfeat), H, W, _ptr(dm[0]), _ptr(dm[1]), _ptr(dm[2]), st), "aarmvs_feat_assemble_bwd")
        for l, (hl, wl) in enumerate(lv):
            pl, ch = hl * wl, (16 if l == 0 else 8)
            groups = max(1, ch // 8)
            # 1x1 head: GroupNorm+ReLU backward, then the conv
            self.gn_bwd(fs[f"m{l}"], dm[l], pl, ch, groups, sl(f"head{l}"), w[f"head{l}.gn"], dm[l], g[f"head{l}.gamma"], g[f"head{l}.beta"],
                        g[f"head{l}.b"], chan=chan, zero=True)
            dcp = torch.empty(pl, 32, **f32)
            self.gn_apply(fs[f"dc{l}"], pl, 32, 4, sl(f"dc{l}"), w[f"dc{l}.gn"], dcp)
            ddc = torch.empty(pl, 32, **f32)
            self.conv([_seg(dm[l], ch)], hl, wl, [_seg(ddc, 32)], hl, wl, 1, 1, True, bw[f"head{l}.wT"])
            self.wgrad([_seg(dcp, 32)], hl, wl, dm[l], hl, wl, ch, 1, 1, g[f"head{l}.w"])
            # deformable conv: GroupNorm+ReLU backward, contraction backward, gather backward
            self.gn_bwd(fs[f"dc{l}"], ddc, pl, 32, 4, sl(f"dc{l}"), w[f"dc{l}.gn"], ddc, g[f"dc{l}.gamma"], g[f"dc{l}.beta"], g[f"dc{l}.b"], chan=chan, zero=True)
            samp, samp_m, dsamp_m = (torch.empty(pl, 288, **f32) for _ in range(3))
            capi.check(lib.aarmvs_deform_sample(_ptr(xp[l]), hl, wl, _ptr(fs[f"om{l}"]), _ptr(samp), _ptr(samp_m), st), "aarmvs_deform_sample")
            self.conv([_seg(ddc, 32)], hl, wl, [_seg(dsamp_m, 288)], hl, wl, 1, 1, False, bw[f"dc{l}.w288T"])
            self.wgrad([_seg(samp_m, 288)], hl, wl, ddc, hl, wl, 32, 1, 1, g[f"dc{l}.w"])
            dom = torch.empty(pl, 32, **f32)
            capi.check(lib.aarmvs_deform_bwd(_ptr(xp[l]), hl, wl, _ptr(fs[f"om{l}"]), _ptr(samp), _ptr(dsamp_m), _ptr(dxp[l]), _ptr(dom), st),
                       "aarmvs_deform_bwd")
            # offset / modulation conv (p_conv | m_conv fused, 27 of 32 output channels used)
            self.conv([_seg(dom, 32)], hl, wl, [_seg(dxp[l], 32, 1)], hl, wl, 3, 1, True, bw[f"pm{l}.wT"])
            self.wgrad([_seg(xp[l], 32)], hl, wl, dom, hl, wl, 32, 3, 1, g[f"pm{l}.w"], g[f"pm{l}.b"])
        # trunk, deepest level first
        for l in (2, 1):
            hl, wl = lv[l]
            self.gn_bwd(fs[f"x{l}"], dxp[l], hl * wl, 32, 4, sl(f"conv{l}"), w[f"conv{l}.gn"], dxp[l], g[f"conv{l}.gamma"], g[f"conv{l}.beta"],
                        g[f"conv{l}.b"], chan=chan, zero=True)
            hi, wi = lv[l - 1]
            self.conv([_seg(dxp[l], 32)], hl, wl, [_seg(dxp[l - 1], 32, 1)], hi, wi, 3, 2, True, bw[f"conv{l}.wT"])
            self.wgrad([_seg(xp[l - 1], 32)], hi, wi, dxp[l], hl, wl, 32, 3, 2, g[f"conv{l}.w"])
        self.gn_bwd(fs["x0"], dxp[0], P, 32, 4, sl("conv0"), w["conv0.gn"], dxp[0], g["conv0.gamma"], g["conv0.beta"], g["conv0.b"], chan=chan, zero=True)
        da2 = torch.empty(P, 16, **f32)
        self.conv([_seg(dxp[0], 32)], H, W, [_seg(da2, 16)], H, W, 3, 1, True, bw["conv0.wT"])
        self.wgrad([_seg(a2p, 16)], H, W, dxp[0], H, W, 32, 3, 1, g["conv0.w"])
        self.gn_bwd(fs["a2"], da2, P, 16, 2, sl("init1"), w["init1.gn"], da2, g["init1.gamma"], g["init1.beta"], g["init1.b"], chan=chan, zero=True)
        da1 = torch.empty(P, 8, **f32)
        self.conv([_seg(da2, 16)], H, W, [_seg(da1, 8)], H, W, 3, 1, True, bw["init1.wT"])
        self.wgrad([_seg(a1p, 8)], H, W, da2, H, W, 16, 3, 1, g["init1.w"])
        self.gn_bwd(fs["a1"], da1, P, 8, 1, sl("init0"), w["init0.gn"], da1, g["init0.gamma"], g["init0.beta"], g["init0.b"], chan=chan, zero=True)
        img4 = torch.zeros(P, 4, **f32)
        img4[:, :3] = img.reshape(3, P).t()
        self.wgrad([_seg(img4, 4)], H, W, da1, H, W, 8, 3, 1, g["init0.w"])

    # ------------------------------------------------------------------------------------------ results
    def parameter_gradients(self):
        """The accumulated gradients in the layout and under the names of the reference's state dict (90 tensors)."""
        g, out = self.g, {}
        cr = "cost_regularization"
        for i in range(5):
            cin, ch = _CELL_IN[i] + _HID[i], _HID[i]
            out[f"{cr}.cell_list.{i}.conv.weight"] = _taps_inverse(g[f"cell{i}.w"], 4 * ch, cin, 3)
            out[f"{cr}.cell_list.{i}.conv.bias"] = g[f"cell{i}.b"].clone()
        for i in range(2):
            out[f"{cr}.deconv_{i}.conv.weight"] = g[f"deconv{i}.wT"].reshape(3, 3, 16, 16).permute(3, 2, 0, 1).contiguous()   # [tap][co][ci] -> [ci][co][ky][kx]
            out[f"{cr}.deconv_{i}.conv.bias"] = g[f"deconv{i}.b"].clone()
            out[f"{cr}.deconv_{i}.gn.weight"] = g[f"deconv{i}.gamma"].clone()
            out[f"{cr}.deconv_{i}.gn.bias"] = g[f"deconv{i}.beta"].clone()
        out[f"{cr}.conv_0.weight"] = _taps_inverse(g["head.w"], 1, 8, 3)
        out[f"{cr}.conv_0.bias"] = g["head.b"].clone()
        om = "omega.reweight_network"
        out[om + ".0.0.weight"] = _taps_inverse(g["stem.w"], 4, 32, 3)
        out[om + ".0.0.bias"] = g["stem.b"].clone()
        t = g["tail"].to(torch.float32)
        s = om + ".1.stem"
        names = [(om + ".0.1.weight", (4,)), (om + ".0.1.bias", (4,)), (s + ".0.0.weight", (4, 4, 1, 1)), (s + ".0.0.bias", (4,)),
                 (s + ".0.1.weight", (4,)), (s + ".0.1.bias", (4,)), (s + ".1.weight", (4, 4, 1, 1)), (s + ".1.bias", (4,)),
                 (s + ".2.weight", (4,)), (s + ".2.bias", (4,)), (om + ".2.weight", (1, 4, 1, 1)), (om + ".2.bias", (1,))]
        off = 0
        for name, shape in names:
            n = 1
            for x in shape:
                n *= x
            out[name] = t[off:off + n].reshape(shape).clone()
            off += n
        ft = "feature"
        for name, prefix, cin, cout, k in (("init0", ft + ".init_conv.0", 3, 8, 3), ("init1", ft + ".init_conv.1", 8, 16, 3),
                                           ("conv0", ft + ".conv0", 16, 32, 3), ("conv1", ft + ".conv1", 32, 32, 3),
                                           ("conv2", ft + ".conv2", 32, 32, 3), ("head0", ft + ".intraAA.conv0", 32, 16, 1),
                                           ("head1", ft + ".intraAA.conv1", 32, 8, 1), ("head2", ft + ".intraAA.conv2", 32, 8, 1)):
            out[prefix + ".0.weight"] = _taps_inverse(g[name + ".w"], cout, cin, k)
            out[prefix + ".0.bias"] = g[name + ".b"].clone()
            out[prefix + ".1.weight"] = g[name + ".gamma"].clone()
            out[prefix + ".1.bias"] = g[name + ".beta"].clone()
        for l in range(3):
            dc = f"{ft}.intraAA.deformconv{l}"
            out[dc + ".0.conv.weight"] = _taps_inverse(g[f"dc{l}.w"].reshape(9, 32, 32), 32, 32, 3)
            out[dc + ".0.conv.bias"] = g[f"dc{l}.b"].clone()
            pm = _taps_inverse(g[f"pm{l}.w"], 27, 32, 3)
            out[dc + ".0.p_conv.weight"], out[dc + ".0.m_conv.weight"] = pm[:18].contiguous(), pm[18:27].contiguous()
            out[dc + ".0.p_conv.bias"], out[dc + ".0.m_conv.bias"] = g[f"pm{l}.b"][:18].clone(), g[f"pm{l}.b"][18:27].clone()
            out[dc + ".1.weight"] = g[f"dc{l}.gamma"].clone()
            out[dc + ".1.bias"] = g[f"dc{l}.beta"].clone()
        return out
