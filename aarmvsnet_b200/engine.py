"""Host-side driver of the sm_100a kernels: owns the device workspace (NHWC fp32, sized once per shape), the packed
weights, and the per-plane kernel chain, which is captured once into a CUDA graph and replayed for every depth
plane (the plane index lives in device memory).  PyTorch is used for device memory and streams only.

Per view (FeatNet, models/drmvsnet.py:41-63):   16 launches.
Per batch of K depth planes (models/drmvsnet.py:306-334): cost_stem, omega_stats x2, aggregate once for the K planes,
                                                then per plane cell0, cell1, cell2, deconv_0, cell3, deconv_1, cell4,
                                                head_wta, next_plane: 4 + 9 K launches, 0 host syncs
                                                (the reference issues ~526 ATen ops and 12 host syncs PER PLANE).
"""
import collections
import ctypes

import torch

from . import capi, pack

_HID = (16, 16, 16, 16, 8)            # hidden channels of the five cells (models/drmvsnet.py:242)
_SCALE = (1, 2, 4, 2, 1)              # their resolution divisors (models/drmvsnet.py:91)


def _ptr(t, offset_elems=0):
    return ctypes.c_void_p(t.data_ptr() + offset_elems * t.element_size())


def _src(t, channels, mode=capi.SRC_PLAIN, *, planes=0, sums=None, gn=None, groups=0, count=0.0, pp_stride=0, pp_phase=0):
    s = capi.ConvSrc()
    s.ptr = t.data_ptr()
    s.channels, s.mode, s.planes = channels, mode, planes
    s.gn_groups = groups
    s.gn_sums = sums.data_ptr() if sums is not None else None
    s.gn_weight = gn[0].data_ptr() if gn is not None else None
    s.gn_bias = gn[1].data_ptr() if gn is not None else None
    s.gn_count = float(count)
    s.pp_stride, s.pp_phase = pp_stride, pp_phase
    return s


def _src_array(*srcs):
    arr = (capi.ConvSrc * len(srcs))(*srcs)
    return arr


def _psrc(t, channels, pp=1, phase=0):
    """Operand planes [pp][2][channels/8][H][W][8] fp16 as a source of the tensor-core cell (struct aarmvs_plane_src)."""
    s = capi.PlaneSrc()
    s.planes = t.data_ptr()
    s.channels, s.pp, s.pp_phase = channels, pp, phase
    return s


def _psrc_array(*srcs):
    return (capi.PlaneSrc * len(srcs))(*srcs)


class Engine:
    """Runs EMVSNet.forward's arithmetic for one model instance (weights are re-packed when they change)."""

    MAX_WORKSPACES = 2        # per-shape workspaces (and their captured graphs) kept alive; least recently used is dropped

    def __init__(self, module, align_corners=False, use_graph=True, cell_math="tc", plane_batch=8, keep_costv=False):
        # cell_math "fma" (fp32 CUDA-core kernels) is a precision reference for the tests, not a product path
        if cell_math not in ("tc", "fma"):
            raise ValueError("cell_math must be 'tc' (tcgen05, split-fp16 operands) or 'fma' (fp32 CUDA cores)")
        self.cell_math = cell_math
        # keep_costv=True: pass A also stores the per-view cost (warp - ref)^2 of a plane batch (11.6 GB at cfg2) and pass D streams
        # it back instead of gathering a second time.  Round 1's default; same-box A/B in round 2 (profiles/r02l_*.json): 0.8712 /
        # 0.8697 depth maps/s at 1 751 MHz under the power cap WITH the slab, 0.8736 / 0.8719 at 1 942 MHz WITHOUT it -- the
        # second gather costs about what the 3 GB/plane of extra DRAM traffic costs in clock, so the default is the variant that
        # never lets a per-view cost reach HBM (BASELINE.json north_star (2)).
        self.keep_costv = keep_costv
        if not 1 <= plane_batch <= 16:
            raise ValueError("plane_batch must be in 1..16")
        self.plane_batch = plane_batch            # planes per launch of passes A-D (L1/L2 reuse of the source features)
        self.module = module
        self.align_corners = bool(align_corners)     # grid_sample default under torch >= 1.3 is False
        self.use_graph = use_graph
        self.lib = capi.load()
        self._wkey = None
        self._ws = collections.OrderedDict()          # LRU of sweep workspaces, see _workspace
        self._fs = collections.OrderedDict()          # LRU of FeatNet scratch buffers, see _feat_scratch

    # ------------------------------------------------------------------------------------------ weights
    def _weights(self, device):
        params = list(self.module.parameters())
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in params)
        if key == self._wkey:
            return self.w
        # Inference: re-pack on the HOST and upload each packed tensor once (a few dozen copies instead of ~1 000 tiny device
        # kernels in front of the first depth map).  Training re-packs after every optimizer step: there the tensors are packed
        # on the device (no 90 device-to-host copies, no host arithmetic while the GPU idles) and only the few small layers
        # whose parameters travel as kernel parameter banks are read back.
        pack_dev = torch.device("cpu") if not self.module.training else device
        sd = {k: v.detach().to(device=pack_dev, dtype=torch.float32) for k, v in self.module.state_dict().items()}
        w = {}

        def convgn(name, prefix, cin_pad=None):
            w[name + ".w"] = pack.conv_taps(sd[prefix + ".0.weight"], cin_pad)
            w[name + ".b"] = sd[prefix + ".0.bias"].contiguous()
            w[name + ".gn"] = (sd[prefix + ".1.weight"].contiguous(), sd[prefix + ".1.bias"].contiguous())

        convgn("init0", "feature.init_conv.0", cin_pad=4)
        convgn("init1", "feature.init_conv.1")
        convgn("conv0", "feature.conv0")
        convgn("conv1", "feature.conv1")
        convgn("conv2", "feature.conv2")
        for l in range(3):
            dc = f"feature.intraAA.deformconv{l}"
            pm_w = torch.cat([sd[dc + ".0.p_conv.weight"], sd[dc + ".0.m_conv.weight"]], 0)
            pm_b = torch.cat([sd[dc + ".0.p_conv.bias"], sd[dc + ".0.m_conv.bias"]], 0)
            w[f"pm{l}.w"] = pack.conv_taps(pm_w, cout_pad=32)
            w[f"pm{l}.b"] = pack.pad_vec(pm_b, 32)
            w[f"dc{l}.w"] = pack.conv_taps(sd[dc + ".0.conv.weight"])
            w[f"dc{l}.wtc"] = pack.deform_taps_f16(sd[dc + ".0.conv.weight"])
            w[f"dc{l}.b"] = sd[dc + ".0.conv.bias"].contiguous()
            w[f"dc{l}.gn"] = (sd[dc + ".1.weight"].contiguous(), sd[dc + ".1.bias"].contiguous())
            convgn(f"head{l}", f"feature.intraAA.conv{l}")
        cr = "cost_regularization"
        for i in range(5):
            w[f"cell{i}.w"], w[f"cell{i}.b"] = pack.lstm_gates(sd[f"{cr}.cell_list.{i}.conv.weight"],
                                                                sd[f"{cr}.cell_list.{i}.conv.bias"])
            w[f"cell{i}.wtc"], _ = pack.lstm_gates_f16_steps(sd[f"{cr}.cell_list.{i}.conv.weight"],
                                                             sd[f"{cr}.cell_list.{i}.conv.bias"])
        for i in range(2):
            w[f"deconv{i}.w_host"] = capi.host_floats(pack.deconv_taps(sd[f"{cr}.deconv_{i}.conv.weight"]).reshape(-1).cpu().tolist())
            w[f"deconv{i}.b_host"] = capi.host_floats(sd[f"{cr}.deconv_{i}.conv.bias"].cpu().tolist())
            w[f"deconv{i}.wtc"] = pack.deconv_pairs_f16(sd[f"{cr}.deconv_{i}.conv.weight"])
            w[f"deconv{i}.gn"] = (sd[f"{cr}.deconv_{i}.gn.weight"].contiguous(), sd[f"{cr}.deconv_{i}.gn.bias"].contiguous())
        # small layers whose parameters are baked into kernel parameter banks (host arrays)
        om = "omega.reweight_network"
        w["stem.w_host"] = capi.host_floats(pack.conv_taps(sd[om + ".0.0.weight"]).reshape(-1).cpu().tolist())
        w["stem.b_host"] = capi.host_floats(sd[om + ".0.0.bias"].cpu().tolist())
        w["stem.wtc"] = pack.stem_taps_f16(sd[om + ".0.0.weight"])
        tail = pack.omega_tail(sd)
        assert tail.numel() == self.lib.aarmvs_omega_tail_floats()
        w["tail_host"] = capi.host_floats(tail.tolist())
        w["head.w_host"] = capi.host_floats(pack.conv_taps(sd[cr + ".conv_0.weight"]).reshape(-1).cpu().tolist())
        w["head.wtc"] = pack.head_pairs_f16(sd[cr + ".conv_0.weight"])
        w["head.b_host"] = capi.host_floats(sd[cr + ".conv_0.bias"].cpu().tolist())
        w = {k: (v.to(device) if torch.is_tensor(v) else tuple(t.to(device) for t in v) if isinstance(v, tuple) else v)
             for k, v in w.items()}
        self.w, self._wkey = w, key
        for ws in self._ws.values():     # captured graphs bake host-side parameters: drop them with the old weights (the
            ws["graph"] = None           # buffers stay: a training loop changes the weights every step)
        return w

    # ------------------------------------------------------------------------------------------ workspace
    def _workspace(self, H, W, N, D, device, keep_cost, with_loss=False, with_state=False, slot=0):
        key = (H, W, N, D, str(device), keep_cost, with_loss, with_state, slot)
        ws = self._ws.get(key)
        if ws is not None:
            self._ws.move_to_end(key)
            return ws
        # a scan mixes view counts / shapes: do not let slabs pile up.  The per-sample slots of one shape (training batches)
        # count as one workspace and leave together.
        bases = {k[:-1] for k in self._ws}
        while key[:-1] not in bases and len(bases) >= self.MAX_WORKSPACES:
            oldest = next(iter(self._ws))[:-1]
            for k in [k for k in self._ws if k[:-1] == oldest]:
                del self._ws[k]
            bases.discard(oldest)
        if H % 4 or W % 4:
            raise RuntimeError(f"feature height and width must be multiples of 4, got {H}x{W} (models/drmvsnet.py:51-52, 148-160)")
        if not 2 <= N <= 16:
            raise RuntimeError(f"2..16 views supported, got {N}")
        f32 = dict(device=device, dtype=torch.float32)
        P, S = H * W, N - 1
        K = min(self.plane_batch, D)
        ws = {"H": H, "W": W, "N": N, "D": D, "K": K, "graph": None, "warm": False}
        ws["feat"] = self.feature_buffer(N, H, W, device)
        ws.update(self._feat_scratch(H, W, device))
        # plane sweep
        ws["t0"] = torch.empty(K, S, P, 4, **f32)
        tc = self.cell_math == "tc"
        # optional: (warp - ref)^2 kept by pass A so that pass D streams instead of gathering again (keep_costv, see __init__)
        ws["costv"] = torch.empty(K, S, P, 32, **f32) if tc and self.keep_costv else None
        ws["cost"] = torch.empty(K, P, 32, **f32) if not tc else None
        if tc:
            # everything the cells read is kept as fp16 (hi, lo) operand planes [2][C/8][H][W][8] (include/aarmvs.h):
            # cost (pass D), h ping-pong pairs and max-pooled h (cell epilogues), GroupNorm+ReLU'd deconv outputs
            f16 = dict(device=device, dtype=torch.float16)
            ws["costp"] = torch.empty(K, 2, 4, P, 8, **f16)
            ws["hp"] = [torch.zeros(2, 2, (ch + 7) // 8, P // (s * s), 8, **f16) for ch, s in zip(_HID, _SCALE)]
            ws["pool"] = [torch.empty(2, 2, P // 4, 8, **f16), torch.empty(2, 2, P // 16, 8, **f16)]
            ws["up0p"] = torch.empty(2, 2, P // 4, 8, **f16)
            ws["up1p"] = torch.empty(2, 2, P, 8, **f16)
            ws["c"] = [torch.empty(self.lib.aarmvs_cell_state_floats(H // s, W // s, ch), **f32) for ch, s in zip(_HID, _SCALE)]
            # no fp32 copies of h on the inference path (deconvolutions and the head read the planes too); the training forward
            # asks the cells for them (h_f32_pair) because the backward pass snapshots the recurrent state of every plane
            ws["h"] = ([torch.empty(2, P // (s * s), ch, **f32) for ch, s in zip(_HID, _SCALE)] if with_state else [None] * 5)
        else:
            ws["h"] = [torch.empty(2, P // (s * s), ch, **f32) for ch, s in zip(_HID, _SCALE)]
            ws["c"] = [torch.empty(P // (s * s), ch, **f32) for ch, s in zip(_HID, _SCALE)]
        if not tc:               # fp32 deconvolution outputs (the tensor-core path writes operand planes directly)
            ws["up0"] = torch.empty(P // 4, 16, **f32)
            ws["up1"] = torch.empty(P, 16, **f32)
        if tc and with_state:
            # training forward: where aarmvs_snapshot_state finds the cells' state, and the device cell holding the base address
            # of the current sample's state history (so that one captured graph serves every sample)
            ws["snap"] = {"h": (ctypes.c_void_p * 5)(*[t.data_ptr() for t in ws["h"]]),
                          "c": (ctypes.c_void_p * 5)(*[t.data_ptr() for t in ws["c"]]),
                          "state_pp": torch.zeros(1, device=device, dtype=torch.int64), "rec": self.state_layout(H, W)[1],
                          # per sample: base addresses of the five cells' gate histories (forward_saving_state)
                          "gates_pp": torch.zeros(5, device=device, dtype=torch.int64)}
        ws["psums"] = torch.zeros(K * S * 6 + 8, device=device, dtype=torch.float64)
        ws["plane"] = torch.zeros(1, device=device, dtype=torch.int32)
        ws["cams"] = torch.empty(S, 12, **f32)
        ws["depths"] = torch.empty(D, **f32)
        ws["max_prob"] = torch.empty(P, **f32)
        ws["exp_sum"] = torch.empty(P, **f32)
        ws["depth"] = torch.empty(P, **f32)
        ws["conf"] = torch.empty(P, **f32)
        ws["index"] = torch.empty(P, device=device, dtype=torch.int32)
        ws["cost_volume"] = torch.empty(D, P, **f32) if keep_cost else None
        # fused training loss (tensor-core path): nearest-hypothesis index, mask, cost at that plane, loss accumulator
        ws["gt"] = ((torch.zeros(P, device=device, dtype=torch.int32), torch.zeros(P, **f32), torch.zeros(3, P, **f32),
                     torch.zeros(1, device=device, dtype=torch.float64)) if with_loss else None)
        self._ws[key] = ws
        return ws

    @staticmethod
    def feature_buffer(N, H, W, device):
        """[N][H*W][32] fp32 feature maps followed by (W + 1) zeroed pixels: the gather kernels read a sample's four lines at
        fixed steps from the north-west one, so a zero-weight east / south corner of the last map's last row may touch up to
        (W + 1) * 128 bytes behind it (include/aarmvs.h: aarmvs_cost_stem_tc, aarmvs_aggregate)."""
        flat = torch.zeros((N * H * W + W + 1) * 32, device=device, dtype=torch.float32)
        return flat[: N * H * W * 32].view(N, H * W, 32)

    def _feat_scratch(self, H, W, device, slot=0):
        """FeatNet scratch (reused for every view), separate from the sweep workspace so features() allocates only this.
        slot: the samples of a training batch run FeatNet's backward (which recomputes its forward into this scratch) on
        concurrent streams, each with its own copy; the slots of one shape count as one entry of the LRU."""
        key = (H, W, str(device), slot)
        fs = self._fs.get(key)
        if fs is not None:
            self._fs.move_to_end(key)
            return fs
        if H % 4 or W % 4:
            raise RuntimeError(f"feature height and width must be multiples of 4, got {H}x{W} (models/drmvsnet.py:51-52, 148-160)")
        bases = {k[:-1] for k in self._fs}
        while key[:-1] not in bases and len(bases) >= self.MAX_WORKSPACES:
            oldest = next(iter(self._fs))[:-1]
            for k in [k for k in self._fs if k[:-1] == oldest]:
                del self._fs[k]
            bases.discard(oldest)
        f32 = dict(device=device, dtype=torch.float32)
        P = H * W
        fs = {"H": H, "W": W, "a1": torch.empty(P, 8, **f32), "a2": torch.empty(P, 16, **f32)}
        for l, s in enumerate((1, 2, 4)):
            pl = P // (s * s)
            fs[f"x{l}"] = torch.empty(pl, 32, **f32)
            fs[f"om{l}"] = torch.empty(pl, 32, **f32)
            fs[f"dc{l}"] = torch.empty(pl, 32, **f32)
            fs[f"m{l}"] = torch.empty(pl, 16 if l == 0 else 8, **f32)
        fs["fsums"] = torch.zeros(64, device=device, dtype=torch.float64)
        self._fs[key] = fs
        return fs

    # ------------------------------------------------------------------------------------------ FeatNet
    def feature_map(self, img, out, ws, w, stream):
        """img: [3,H,W] fp32 contiguous (NCHW plane order); out: [P,32] NHWC feature map."""
        lib, H, W = self.lib, ws["H"], ws["W"]
        sums = ws["fsums"]
        sums.zero_()
        off = {"init0": 0, "init1": 2, "conv0": 6, "conv1": 14, "conv2": 22, "dc0": 30, "dc1": 38, "dc2": 46,
               "head0": 54, "head1": 58, "head2": 60}

        def sl(name):
            return sums[off[name]:]

        def conv(name, srcs, hin, win, cin, cout, k, stride, dst, stored, stat):
            arr = _src_array(*srcs)
            capi.check(lib.aarmvs_conv2d(arr, len(srcs), hin, win, cin, cout, k, stride, _ptr(w[name + ".w"]),
                                         _ptr(w[name + ".b"]), _ptr(dst), stored,
                                         _ptr(sl(stat)) if stat else None, stream), "aarmvs_conv2d:" + name)

        def gn(t, ch, name, wname, npix):
            groups = max(1, ch // 8)
            return _src(t, ch, capi.SRC_GN_RELU, sums=sl(name), gn=w[wname + ".gn"], groups=groups,
                        count=(ch // groups) * npix)

        P = H * W
        conv("init0", [_src(img, 4, capi.SRC_NCHW, planes=3)], H, W, 4, 8, 3, 1, ws["a1"], 8, "init0")
        conv("init1", [gn(ws["a1"], 8, "init0", "init0", P)], H, W, 8, 16, 3, 1, ws["a2"], 16, "init1")
        conv("conv0", [gn(ws["a2"], 16, "init1", "init1", P)], H, W, 16, 32, 3, 1, ws["x0"], 32, "conv0")
        conv("conv1", [gn(ws["x0"], 32, "conv0", "conv0", P)], H, W, 32, 32, 3, 2, ws["x1"], 32, "conv1")
        conv("conv2", [gn(ws["x1"], 32, "conv1", "conv1", P // 4)], H // 2, W // 2, 32, 32, 3, 2, ws["x2"], 32, "conv2")
        for l, s in enumerate((1, 2, 4)):
            hl, wl, pl = H // s, W // s, P // (s * s)
            x = gn(ws[f"x{l}"], 32, f"conv{l}", f"conv{l}", pl)
            conv(f"pm{l}", [x], hl, wl, 32, 32, 3, 1, ws[f"om{l}"], 32, None)
            if self.cell_math == "tc":
                capi.check(lib.aarmvs_deform_conv_tc(ctypes.byref(x), hl, wl, _ptr(ws[f"om{l}"]), _ptr(w[f"dc{l}.wtc"]),
                                                     _ptr(w[f"dc{l}.b"]), _ptr(ws[f"dc{l}"]), _ptr(sl(f"dc{l}")), stream),
                           "aarmvs_deform_conv_tc")
            else:
                capi.check(lib.aarmvs_deform_conv(ctypes.byref(x), hl, wl, _ptr(ws[f"om{l}"]), _ptr(w[f"dc{l}.w"]),
                                                  _ptr(w[f"dc{l}.b"]), _ptr(ws[f"dc{l}"]), _ptr(sl(f"dc{l}")), stream),
                           "aarmvs_deform_conv")
            ch = 16 if l == 0 else 8
            conv(f"head{l}", [gn(ws[f"dc{l}"], 32, f"dc{l}", f"dc{l}", pl)], hl, wl, 32, ch, 1, 1, ws[f"m{l}"], ch,
                 f"head{l}")
        g0, g1, g2 = w["head0.gn"], w["head1.gn"], w["head2.gn"]
        capi.check(lib.aarmvs_feat_assemble(_ptr(ws["m0"]), _ptr(ws["m1"]), _ptr(ws["m2"]), _ptr(sl("head0")),
                                            _ptr(sl("head1")), _ptr(sl("head2")), _ptr(g0[0]), _ptr(g0[1]),
                                            _ptr(g1[0]), _ptr(g1[1]), _ptr(g2[0]), _ptr(g2[1]), H, W, _ptr(out), stream),
                   "aarmvs_feat_assemble")

    def cost_stem(self, ref, src, view_stride, H, W, S, cams, w, depths, d_host, dptr, planes, t0, sums, stream, costv=None):
        """Pass A (gather, (warp - ref)^2, omega stem conv) on the tensor cores ('tc'; with costv the per-view cost is also
        stored so that pass D streams it instead of gathering again) or the fp32 FMA pipe ('fma', test reference)."""
        ac = int(self.align_corners)
        if self.cell_math == "tc":
            capi.check(self.lib.aarmvs_cost_stem_tc(ref, src, view_stride, H, W, S, cams, _ptr(w["stem.wtc"]), w["stem.b_host"],
                                                    depths, d_host, dptr, planes, ac, t0, sums, costv, 1 if costv is not None else 0,
                                                    stream), "aarmvs_cost_stem_tc")
        else:
            capi.check(self.lib.aarmvs_cost_stem(ref, src, view_stride, H, W, S, cams, w["stem.w_host"], w["stem.b_host"],
                                                 depths, d_host, dptr, planes, ac, t0, sums, stream), "aarmvs_cost_stem")

    def split_planes(self, src, H, W, planes, stream):
        """fp32 NHWC source (optionally GroupNorm+ReLU'd on the fly) -> fp16 (hi, lo) operand planes of the tc cells."""
        capi.check(self.lib.aarmvs_split_planes(ctypes.byref(src), H, W, _ptr(planes), stream), "aarmvs_split_planes")

    # ------------------------------------------------------------------------------------------ one plane
    def _batch(self, ws, w, stream, nplanes, marks=None):
        """Enqueue the kernel chain of `nplanes` consecutive depth planes starting at the plane index held in
        ws['plane'] (device memory): passes A-D once for the whole batch, then the recurrence plane by plane.
        marks: optional list collecting (stage name, CUDA event) after each stage (bench.py's per-kernel timing)."""
        lib, H, W, N = self.lib, ws["H"], ws["W"], ws["N"]

        def mark(name):
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(torch.cuda.current_stream())
                marks.append((name, ev))

        mark("begin")
        P, S = H * W, N - 1
        feat, dptr, ac = ws["feat"], _ptr(ws["plane"]), int(self.align_corners)
        ref, src = _ptr(feat), _ptr(feat, P * 32)
        ps = ws["psums"]
        n_omega = ws["K"] * S * 6
        costv = _ptr(ws["costv"]) if ws["costv"] is not None else None
        self.cost_stem(ref, src, P * 32, H, W, S, _ptr(ws["cams"]), w, _ptr(ws["depths"]), 0, dptr, nplanes, _ptr(ws["t0"]),
                       _ptr(ps), stream, costv)
        mark("cost_stem")
        for stage in (2, 3):
            capi.check(lib.aarmvs_omega_stats(stage, _ptr(ws["t0"]), H, W, S, nplanes, w["tail_host"], _ptr(ps), stream),
                       "aarmvs_omega_stats")
        mark("omega_stats")
        tc = self.cell_math == "tc"
        # tensor-core cells: pass D (streaming or gathering) writes the cost straight into cell 0's operand planes
        direct = tc
        capi.check(lib.aarmvs_aggregate(ref, src, P * 32, H, W, S, _ptr(ws["cams"]), w["tail_host"], _ptr(ws["depths"]), 0,
                                        dptr, nplanes, ac, _ptr(ws["t0"]), _ptr(ps), costv,
                                        None if direct else _ptr(ws["cost"]), _ptr(ws["costp"]) if direct else None, None,
                                        stream), "aarmvs_aggregate")
        mark("aggregate")
        h, c = ws["h"], ws["c"]

        def hsrc(i, mode=capi.SRC_PLAIN):          # h_i of the CURRENT plane (ping-pong half d & 1)
            return _src(h[i], _HID[i], mode, pp_stride=h[i][0].numel(), pp_phase=0)

        def cell(i, srcs, hh, ww, cin):
            arr = _src_array(*srcs)
            capi.check(lib.aarmvs_convlstm_cell(arr, len(srcs), hh, ww, cin, _HID[i], _ptr(w[f"cell{i}.w"]),
                                                _ptr(w[f"cell{i}.b"]), _ptr(h[i]), _ptr(c[i]), 0, dptr, stream),
                       f"aarmvs_convlstm_cell[{i}]")
            mark(f"cell{i}")

        def cell_tc(i, srcs, hh, ww, cin, pooled=None):
            arr = _psrc_array(*srcs)
            capi.check(lib.aarmvs_convlstm_cell_tc(arr, len(srcs), hh, ww, cin, _HID[i], _ptr(w[f"cell{i}.wtc"]),
                                                   _ptr(w[f"cell{i}.b"]), _ptr(ws["hp"][i]), _ptr(c[i]),
                                                   _ptr(h[i]) if h[i] is not None else None,
                                                   _ptr(pooled) if pooled is not None else None,
                                                   _ptr(ws["snap"]["gates_pp"], i) if ws.get("snap") is not None else None,
                                                   0, dptr, stream),
                       f"aarmvs_convlstm_cell_tc[{i}]")
            mark(f"cell{i}")

        def hplanes(i):          # h_i of the CURRENT plane as operand planes (ping-pong half d & 1)
            return _psrc(ws["hp"][i], _HID[i], pp=2, phase=0)

        def deconv(i, src_, hin, win, dst, sums):
            capi.check(lib.aarmvs_deconv2x(ctypes.byref(src_), hin, win, w[f"deconv{i}.w_host"], w[f"deconv{i}.b_host"],
                                           _ptr(dst), _ptr(sums), 0, dptr, stream), f"aarmvs_deconv2x[{i}]")
            mark(f"deconv{i}")

        def deconv_tc(i, src_, hin, win, sums, out_planes):
            """deConvGnReLU on the tensor cores: statistics pass, then the same GEMM again with GroupNorm + ReLU + split."""
            g = w[f"deconv{i}.gn"]
            for out in (None, out_planes):
                capi.check(lib.aarmvs_deconv2x_tc(ctypes.byref(src_), hin, win, _ptr(w[f"deconv{i}.wtc"]), w[f"deconv{i}.b_host"],
                                                  _ptr(sums), _ptr(g[0]), _ptr(g[1]), _ptr(out) if out is not None else None,
                                                  0, dptr, stream), f"aarmvs_deconv2x_tc[{i}]")
            mark(f"deconv{i}")

        H2, W2, H4, W4 = H // 2, W // 2, H // 4, W // 4
        s_up0, s_up1 = ps[n_omega:], ps[n_omega + 4:]
        cv = ws["cost_volume"]
        for k in range(nplanes):
            if tc:
                cell_tc(0, [_psrc(ws["costp"][k], 32)], H, W, 48, pooled=ws["pool"][0])
                cell_tc(1, [_psrc(ws["pool"][0], 16)], H2, W2, 32, pooled=ws["pool"][1])
                cell_tc(2, [_psrc(ws["pool"][1], 16)], H4, W4, 32)
                deconv_tc(0, hplanes(2), H4, W4, s_up0, ws["up0p"])
                cell_tc(3, [_psrc(ws["up0p"], 16), hplanes(1)], H2, W2, 48)
                deconv_tc(1, hplanes(3), H2, W2, s_up1, ws["up1p"])
                cell_tc(4, [_psrc(ws["up1p"], 16), hplanes(0)], H, W, 40)
            else:
                up0 = _src(ws["up0"], 16, capi.SRC_GN_RELU, sums=s_up0, gn=w["deconv0.gn"], groups=2, count=8 * (P // 4))
                up1 = _src(ws["up1"], 16, capi.SRC_GN_RELU, sums=s_up1, gn=w["deconv1.gn"], groups=2, count=8 * P)
                cell(0, [_src(ws["cost"][k], 32)], H, W, 48)
                cell(1, [hsrc(0, capi.SRC_MAXPOOL2)], H2, W2, 32)
                cell(2, [hsrc(1, capi.SRC_MAXPOOL2)], H4, W4, 32)
                deconv(0, hsrc(2), H4, W4, ws["up0"], s_up0)
                cell(3, [up0, hsrc(1)], H2, W2, 48)
                deconv(1, hsrc(3), H2, W2, ws["up1"], s_up1)
                cell(4, [up1, hsrc(0)], H, W, 40)
            if tc:
                capi.check(lib.aarmvs_head_wta_tc(_ptr(ws["hp"][4]), H, W, _ptr(w["head.wtc"]), w["head.b_host"], _ptr(ws["depths"]),
                                                  ws["D"], 0, dptr, _ptr(ws["max_prob"]), _ptr(ws["depth"]), _ptr(ws["exp_sum"]),
                                                  _ptr(ws["index"]), _ptr(ws["conf"]), _ptr(cv) if cv is not None else None,
                                                  *([_ptr(t) for t in ws["gt"]] if ws["gt"] is not None else [None] * 4),
                                                  stream), "aarmvs_head_wta_tc")
            else:
                h4 = hsrc(4)
                capi.check(lib.aarmvs_head_wta(ctypes.byref(h4), H, W, w["head.w_host"], w["head.b_host"], _ptr(ws["depths"]),
                                               ws["D"], 0, dptr, _ptr(ws["max_prob"]), _ptr(ws["depth"]), _ptr(ws["exp_sum"]),
                                               _ptr(ws["index"]), _ptr(ws["conf"]), _ptr(cv) if cv is not None else None, stream),
                           "aarmvs_head_wta")
            # ++plane; zero the deconv GroupNorm sums (and, after the last plane of the batch, the omega sums too)
            if k == nplanes - 1:
                capi.check(lib.aarmvs_next_plane(dptr, _ptr(ps), ps.numel(), stream), "aarmvs_next_plane")
            else:
                capi.check(lib.aarmvs_next_plane(dptr, _ptr(s_up0), 8, stream), "aarmvs_next_plane")
            mark("head_wta")
            if ws.get("snap") is not None:          # training forward: record the recurrent state after this plane
                sn = ws["snap"]
                capi.check(lib.aarmvs_snapshot_state(sn["h"], sn["c"], H, W, _ptr(sn["state_pp"]), sn["rec"], 0, dptr, stream),
                           "aarmvs_snapshot_state")

    def _capture_batch(self, ws, w, device, nplanes):
        """Capture the kernel chain of one plane batch as a CUDA graph (raw capture API on a side stream: nothing is allocated
        inside, and torch.cuda.graph()'s context manager would flush the caching allocator and synchronise the device, which a
        training loop -- one capture per optimisation step, the graph bakes host-side weights -- cannot afford)."""
        cur = torch.cuda.current_stream(device)
        side = torch.cuda.Stream(device)
        side.wait_stream(cur)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            g.capture_begin()
            self._batch(ws, w, ctypes.c_void_p(side.cuda_stream), nplanes)
            g.capture_end()
        cur.wait_stream(side)
        return g

    def _run_planes(self, ws, w, device, stream, D):
        """All D planes of one sample: whole batches of K planes through the (lazily captured) graph, the rest eagerly."""
        K = ws["K"]
        full, rem = divmod(D, K)
        start = 0
        if self.use_graph and ws["graph"] is None and full > 0:
            self._batch(ws, w, stream, K)         # first batch eagerly: also sets the kernels' launch attributes
            start = 1
            ws["graph"] = self._capture_batch(ws, w, device, K)
        for _ in range(start, full):
            if self.use_graph:
                ws["graph"].replay()
            else:
                self._batch(ws, w, stream, K)
        if rem:
            self._batch(ws, w, stream, rem)       # last partial batch

    # ------------------------------------------------------------------------------------------ one depth map
    @staticmethod
    def relative_cameras(proj_matrices):
        """rot | trans of src_proj @ inverse(ref_proj) per source view (models/module.py:16-18), evaluated on the host
        with the same torch ops as the reference so the fp32 values are the reference's.  [N,4,4] -> [N-1,12]."""
        pm = proj_matrices.detach().to("cpu", torch.float32)        # no copy / sync when the caller already holds host matrices
        ref = pm[0:1]
        proj = torch.matmul(pm[1:], torch.inverse(ref))
        return torch.cat([proj[:, :3, :3].reshape(-1, 9), proj[:, :3, 3]], 1).contiguous()

    def features(self, img):
        """FeatNet of one image [3,H,W] -> a new [H*W, 32] NHWC feature tensor (for per-scan feature caching)."""
        device = img.device
        if device.type != "cuda":
            raise RuntimeError("aarmvsnet_b200 has no CPU path: inputs must live on a CUDA device")
        _, H, W = img.shape
        with torch.cuda.device(device):
            w = self._weights(device)
            fs = self._feat_scratch(H, W, device)
            out = torch.empty(H * W, 32, device=device, dtype=torch.float32)
            stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
            self.feature_map(img.to(torch.float32).contiguous(), out, fs, w, stream)
            return out

    @staticmethod
    def state_layout(H, W):
        """Offsets (floats) of h_i and c_i, i = 0..4, inside one per-plane state record [h0 | c0 | h1 | c1 | ...], NHWC fp32,
        and the record length: 66 floats per full-resolution pixel."""
        off, lay = 0, []
        for ch, s in zip(_HID, _SCALE):
            n = (H // s) * (W // s) * ch
            lay.append((off, off + n, n))
            off += 2 * n
        return lay, off

    def sweep_streams(self, device, n):
        """n persistent streams of `device` for the concurrent plane sweeps of a training batch (forward_saving_state)."""
        if not hasattr(self, "_sweep_streams"):
            self._sweep_streams = {}
        pool = self._sweep_streams.setdefault(str(device), [])
        while len(pool) < n:
            pool.append(torch.cuda.Stream(device))
        return pool[:n]

    def forward_saving_state(self, imgs, proj_matrices, depth_values, slot=0, sweep_stream=None):
        """The training-phase forward of ONE sample (models/drmvsnet.py:272-287) on the tensor-core kernels (the same graph-replayed
        plane chain as inference, plus aarmvs_snapshot_state after every plane), keeping what the backward pass needs: the features
        of the N views, the regularised cost of every plane, a snapshot of the recurrent state (h, c of the five cells) after every
        plane -- one flat record of 66 floats per full-resolution pixel and plane (state_layout), record 0 = the zero state before
        plane 0 -- and the activated gates of every cell and plane (132 floats per pixel and plane); the rest is recomputed per
        plane by the backward (aarmvsnet_b200/backward.py).
        slot / sweep_stream: the samples of a batch are independent chains of short kernels on small maps, so a caller may run
        each sample's plane sweep on its own stream with its own workspace (`slot`); FeatNet (shared scratch) and every
        allocation stay on the calling stream, and the CALLER joins `sweep_stream` before it touches the results.
        Returns {"cost": [D,P], "feat": [N,P,32], "state": [D+1, record], "gates": 5 x [D,P_i,4 Ch_i], ...}."""
        device = imgs.device
        if device.type != "cuda":
            raise RuntimeError("aarmvsnet_b200 has no CPU path: inputs must live on a CUDA device")
        if self.cell_math != "tc":
            raise RuntimeError("the training forward runs on the tensor-core kernels (cell_math='tc')")
        with torch.cuda.device(device):
            N, _, H, W = imgs.shape
            D = depth_values.shape[0]
            w = self._weights(device)
            ws = self._workspace(H, W, N, D, device, True, False, True, slot)
            main = torch.cuda.current_stream(device)
            stream = ctypes.c_void_p(main.cuda_stream)
            imgs = imgs.to(torch.float32).contiguous()
            for v in range(N):
                self.feature_map(imgs[v], ws["feat"][v], ws, w, stream)
            lay, rec = self.state_layout(H, W)
            f32 = dict(device=device, dtype=torch.float32)
            state = torch.empty(D + 1, rec, **f32)
            state[0].zero_()
            cams_host = self.relative_cameras(proj_matrices)
            saved = {"H": H, "W": W, "N": N, "D": D, "imgs": imgs, "state": state,
                     "cams": torch.empty_like(ws["cams"]), "depths": torch.empty_like(ws["depths"]),
                     "feat": torch.empty_like(ws["feat"]), "cost": torch.empty_like(ws["cost_volume"]),
                     # the activated gates of every plane, as the cells used them: [D][P_i][4 Ch_i] per cell; the kernels read
                     # the bases from device memory, so the captured plane graph serves every sample
                     "gates": [torch.empty(D, (H // sc) * (W // sc), 4 * ch, **f32) for ch, sc in zip(_HID, _SCALE)]}
            dv32 = depth_values.to(torch.float32)
            if sweep_stream is not None:
                sweep_stream.wait_stream(main)
                if dv32 is not depth_values:                 # a temporary of the calling stream, read on the sweep stream
                    dv32.record_stream(sweep_stream)
            depth_values = dv32
            with torch.cuda.stream(sweep_stream if sweep_stream is not None else main):
                stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
                ws["cams"].copy_(cams_host, non_blocking=True)
                ws["depths"].copy_(depth_values, non_blocking=True)
                ws["plane"].zero_()
                ws["psums"].zero_()
                ws["snap"]["state_pp"].fill_(state.data_ptr())
                for i, g in enumerate(saved["gates"]):
                    ws["snap"]["gates_pp"][i].fill_(g.data_ptr())
                self._run_planes(ws, w, device, stream, D)
                for k in ("cams", "depths", "feat"):
                    saved[k].copy_(ws[k])
                saved["cost"].copy_(ws["cost_volume"])
            return saved

    @staticmethod
    def nearest_hypothesis(depth_values, depth_gt, mask):
        """gt_index of mvsnet_cls_loss (models/drmvsnet.py:357-362): argmin_d |depth_values[d] - depth_gt| (first minimum),
        zeroed where the mask is 0 -- from the two neighbouring hypotheses of a monotone sampling instead of a [D,H,W]
        volume.  depth_values [D], depth_gt / mask [H,W] -> int32 [H*W]."""
        dv = depth_values.to(torch.float32)
        D = dv.numel()
        inc = bool(dv[-1] >= dv[0])
        srt = dv if inc else dv.flip(0)
        gt = depth_gt.to(torch.float32).reshape(-1)
        hi = torch.bucketize(gt, srt).clamp(0, D - 1)                  # first sorted hypothesis >= gt
        lo = (hi - 1).clamp(0, D - 1)
        if not inc:
            hi, lo = D - 1 - hi, D - 1 - lo
        first, second = torch.minimum(lo, hi), torch.maximum(lo, hi)  # argmin keeps the lower index on ties
        idx = torch.where((dv[second] - gt).abs() < (dv[first] - gt).abs(), second, first)
        return torch.round(mask.to(torch.float32).reshape(-1) * idx.to(torch.float32)).to(torch.int32)

    def depth_map(self, imgs, proj_matrices, depth_values, keep_cost=False, features=None, loss_gt=None):
        """imgs [N,3,H,W], proj_matrices [N,4,4], depth_values [D] -> dict of [H,W] maps (views of the workspace).
        features: optional list of N cached [H*W,32] feature tensors (from features()); FeatNet is then skipped.
        loss_gt: optional (depth_gt [H,W], mask [H,W]): the head also accumulates the masked cross entropy of
        mvsnet_cls_loss -> out["ce_sum"] (float64 scalar tensor), without any [D,H,W] volume (tensor-core path)."""
        device = imgs.device
        if device.type != "cuda":
            raise RuntimeError("aarmvsnet_b200 has no CPU path: inputs must live on a CUDA device")
        with torch.cuda.device(device):          # the C ABI launches on the calling thread's current device
            return self._depth_map(imgs, proj_matrices, depth_values, keep_cost, features, loss_gt)

    def _depth_map(self, imgs, proj_matrices, depth_values, keep_cost, features, loss_gt=None):
        device = imgs.device
        N, _, H, W = imgs.shape
        D = depth_values.shape[0]
        if D < 1:
            raise RuntimeError("depth_values must hold at least one plane")
        w = self._weights(device)
        if loss_gt is not None and self.cell_math != "tc":
            raise RuntimeError("the fused loss lives in the tensor-core head kernel (cell_math='tc')")
        ws = self._workspace(H, W, N, D, device, keep_cost, loss_gt is not None)
        if loss_gt is not None:
            gt_index, gt_mask, _, loss_acc = ws["gt"]
            gt_index.copy_(self.nearest_hypothesis(depth_values.to(device), loss_gt[0].to(device), loss_gt[1].to(device)))
            gt_mask.copy_(loss_gt[1].to(device=device, dtype=torch.float32).reshape(-1))
            loss_acc.zero_()
        stream = ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        if features is not None:
            if len(features) != N:
                raise RuntimeError(f"{N} views but {len(features)} cached feature maps")
            for v in range(N):
                ws["feat"][v].copy_(features[v], non_blocking=True)
        else:
            imgs = imgs.to(torch.float32).contiguous()
            for v in range(N):
                self.feature_map(imgs[v], ws["feat"][v], ws, w, stream)
        ws["cams"].copy_(self.relative_cameras(proj_matrices), non_blocking=True)
        ws["depths"].copy_(depth_values.to(torch.float32), non_blocking=True)
        ws["plane"].zero_()
        ws["psums"].zero_()
        self._run_planes(ws, w, device, stream, D)
        out = {"depth": ws["depth"].view(H, W), "photometric_confidence": ws["conf"].view(H, W),
               "index": ws["index"].view(H, W)}
        if keep_cost:
            out["cost_reg"] = ws["cost_volume"].view(D, H, W)
        if loss_gt is not None:
            out["ce_sum"] = ws["gt"][3][0]
        return out
