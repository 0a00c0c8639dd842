"""Drop-in mirror of the reference's model API for the per-depth-plane cost-volume path.

`EMVSNet` keeps the reference's constructor, `forward(imgs, proj_matrices, depth_values)` signature, output keys
and parameter names (models/drmvsnet.py:234-345), so the shipped checkpoints load with strict=True; the arithmetic
runs in the sm_100a kernels behind include/aarmvs.h.  The submodules below only HOLD parameters under the
reference's names; they are never called layer by layer.

Differences from the reference, all documented in DESIGN.md:
  * the evidential head (evidential/models.py) is out of scope: it has no weights in the shipped checkpoints and
    raises for D != 32, so `evidential_prediction` is None;
  * no `.cuda()` calls are hard-coded: the module follows its inputs' device, and raises if that is not CUDA;
  * the test-phase output has an extra key "index" (winning plane per pixel, int32).
"""
import torch
import torch.nn as nn

from .engine import Engine


def _conv_gn(cin, cout, k, stride=1):
    # parameter names "<prefix>.0.{weight,bias}" (conv) and "<prefix>.1.{weight,bias}" (GroupNorm) as produced by
    # the reference's convgnrelu (models/module.py:98-103); groups = max(1, cout / 8)
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride=stride, padding=(k - 1) // 2),
                         nn.GroupNorm(int(max(1, cout / 8)), cout))


class DeformConv2d(nn.Module):
    """Parameters of the modulated deformable conv (models/module.py:105-126): conv (k3, applied to the 9 gathered
    taps), p_conv (18 offsets), m_conv (9 modulation logits); p_conv / m_conv start at zero like the reference."""

    def __init__(self, inc=32, outc=32):
        super().__init__()
        self.conv = nn.Conv2d(inc, outc, 3, stride=3, bias=True)
        self.p_conv = nn.Conv2d(inc, 18, 3, padding=1)
        self.m_conv = nn.Conv2d(inc, 9, 3, padding=1)
        nn.init.constant_(self.p_conv.weight, 0)
        nn.init.constant_(self.m_conv.weight, 0)


class IntraViewAAModule(nn.Module):
    """models/drmvsnet.py:7-16."""

    def __init__(self):
        super().__init__()
        for l, ch in enumerate((16, 8, 8)):
            setattr(self, f"deformconv{l}", nn.Sequential(DeformConv2d(32, 32), nn.GroupNorm(4, 32)))
            setattr(self, f"conv{l}", _conv_gn(32, ch, 1))


class FeatNet(nn.Module):
    """models/drmvsnet.py:41-53."""

    def __init__(self):
        super().__init__()
        self.init_conv = nn.Sequential(_conv_gn(3, 8, 3), _conv_gn(8, 16, 3))
        self.conv0 = _conv_gn(16, 32, 3)
        self.conv1 = _conv_gn(32, 32, 3, stride=2)
        self.conv2 = _conv_gn(32, 32, 3, stride=2)
        self.intraAA = IntraViewAAModule()


class _ResBlockParams(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.stem = nn.Sequential(_conv_gn(ch, ch, 1), nn.Conv2d(ch, ch, 1), nn.GroupNorm(int(max(1, ch / 8)), ch))


class InterViewAAModule(nn.Module):
    """models/drmvsnet.py:27-35 (omega)."""

    def __init__(self, in_channels=32):
        super().__init__()
        self.reweight_network = nn.Sequential(_conv_gn(in_channels, 4, 3), _ResBlockParams(4), nn.Conv2d(4, 1, 1))


class ConvLSTMCell(nn.Module):
    """models/module.py:40-74."""

    def __init__(self, input_dim, hidden_dim):
        super().__init__()
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self.conv = nn.Conv2d(input_dim + hidden_dim, 4 * hidden_dim, 3, padding=1)


class _DeconvGnParams(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.ConvTranspose2d(16, 16, 3, stride=2, padding=1, output_padding=1)
        self.gn = nn.GroupNorm(2, 16)


class UNetConvLSTM(nn.Module):
    """models/drmvsnet.py:66-117."""

    def __init__(self, input_dim=(32, 16, 16, 32, 32), hidden_dim=(16, 16, 16, 16, 8)):
        super().__init__()
        self.cell_list = nn.ModuleList([ConvLSTMCell(i, h) for i, h in zip(input_dim, hidden_dim)])
        self.deconv_0 = _DeconvGnParams()
        self.deconv_1 = _DeconvGnParams()
        self.conv_0 = nn.Conv2d(8, 1, 3, 1, padding=1)


class _RegularisedCostVolume(torch.autograd.Function):
    """The differentiable training-phase forward (models/drmvsnet.py:272-289): imgs, cameras, depth hypotheses and the 90
    parameters -> regularised cost volume [B,D,H,W] (the softmax over D stays a torch op on top).  forward = the tensor-core
    engine with per-plane state snapshots; backward = aarmvsnet_b200/backward.py (reverse sweep over the planes with per-plane
    recompute).  Gradients go to the parameters only: the reference builds the sampling grid under no_grad (module.py:15-34)
    and images carry no gradient in training."""

    @staticmethod
    def forward(ctx, model, imgs, proj_matrices, depth_values, *params):
        eng = model._engine
        B, _, _, H, W = imgs.shape
        # the samples' plane sweeps are independent, latency-bound chains: one stream and workspace per sample
        main = torch.cuda.current_stream(imgs.device)
        pool = eng.sweep_streams(imgs.device, B)
        streams = pool if B > 1 else [None]
        saved = [eng.forward_saving_state(imgs[b], proj_matrices[b], depth_values[b], slot=b, sweep_stream=streams[b]) for b in range(B)]
        for st in streams:
            if st is not None:
                main.wait_stream(st)
        ctx.model, ctx.saved = model, saved
        return torch.stack([s["cost"].view(-1, H, W) for s in saved])

    @staticmethod
    def backward(ctx, dcost):
        from .backward import Backward
        model, saved = ctx.model, ctx.saved
        bwd = Backward(model._engine, dcost.device, dx_math=getattr(model, "backward_math", "tc"))
        with torch.cuda.device(dcost.device):
            bwd.batch(saved, dcost.reshape(len(saved), saved[0]["D"], -1).to(torch.float32).contiguous())
        grads = bwd.parameter_gradients()
        ctx.saved = None
        return (None, None, None, None) + tuple(grads[name].to(p.dtype) for name, p in model.named_parameters())


class EMVSNet(nn.Module):
    """Same constructor as models/drmvsnet.py:235.  `disparity_level` only parameterised the evidential head and is
    accepted and ignored; `max_h * image_scale` x `max_w * image_scale` must equal the tensor size in the reference
    (it sizes the zero hidden state, drmvsnet.py:239) -- here the state is sized from the input, and a mismatch with
    a non-default max_h/max_w raises the same way the reference's torch.cat would."""

    def __init__(self, disparity_level=32, image_scale=0.25, max_h=960, max_w=480, return_depth=False,
                 align_corners=False, use_graph=True, strict_size=False, cell_math="tc", plane_batch=8, keep_costv=False):
        super().__init__()
        self.feature = FeatNet()
        self.cost_regularization = UNetConvLSTM()
        self.omega = InterViewAAModule(32)
        self.return_depth = return_depth
        self.input_size = (int(max_h * image_scale), int(max_w * image_scale))
        self.strict_size = strict_size
        self._engine = Engine(self, align_corners=align_corners, use_graph=use_graph, cell_math=cell_math,
                              plane_batch=plane_batch, keep_costv=keep_costv)

    def features(self, img):
        """FeatNet of one image [3,H,W] as the kernels' NHWC feature tensor [H*W,32] (cacheable across reference views)."""
        with torch.no_grad():
            return self._engine.features(img)

    def forward(self, imgs, proj_matrices, depth_values):
        if imgs.dim() != 5 or imgs.shape[2] != 3:
            raise RuntimeError(f"imgs must be [B,N,3,H,W], got {tuple(imgs.shape)}")
        B, N, _, H, W = imgs.shape
        if proj_matrices.shape[:2] != (B, N) or tuple(proj_matrices.shape[2:]) != (4, 4):
            raise AssertionError("Different number of images and projection matrices")     # drmvsnet.py:258
        if depth_values.dim() != 2 or depth_values.shape[0] != B:
            raise RuntimeError(f"depth_values must be [B,D], got {tuple(depth_values.shape)}")
        if self.strict_size and (H, W) != self.input_size:
            raise RuntimeError(f"input {H}x{W} does not match max_h*image_scale x max_w*image_scale = {self.input_size}")
        if not self.return_depth and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training phase with autograd (train.py:297-306): loss.backward() reaches every parameter
            cost = _RegularisedCostVolume.apply(self, imgs, proj_matrices, depth_values, *self.parameters())
            return torch.softmax(cost, dim=1), None, None
        keep_cost = not self.return_depth
        outs = []
        with torch.no_grad():
            for b in range(B):
                o = self._engine.depth_map(imgs[b], proj_matrices[b], depth_values[b], keep_cost=keep_cost)
                outs.append({k: v.clone() for k, v in o.items()})
        if self.return_depth:        # test phase (drmvsnet.py:299-345)
            return {"depth": torch.stack([o["depth"] for o in outs]),
                    "photometric_confidence": torch.stack([o["photometric_confidence"] for o in outs]),
                    "index": torch.stack([o["index"] for o in outs]),
                    "evidential_prediction": None}
        # training phase (drmvsnet.py:289-295): softmax over the D regularised cost maps
        prob_volume = torch.softmax(torch.stack([o["cost_reg"] for o in outs]), dim=1)
        return prob_volume, None, None


    def forward_loss(self, imgs, proj_matrices, depth_values, depth_gt, mask):
        """Training-phase forward fused with mvsnet_cls_loss (models/drmvsnet.py:272-295 + 347-381): returns
        (loss, wta_depth_map) = mvsnet_cls_loss(self_train(imgs, ...)[0], depth_gt, mask, depth_values), but the masked
        cross entropy -log softmax(cost)[gt] = log(sum_d exp(cost_d)) - cost_gt is accumulated by the head kernel plane by
        plane, so the [B,D,H,W] probability volume is never stored.  No autograd."""
        B = imgs.shape[0]
        losses, depths = [], []
        with torch.no_grad():
            for b in range(B):
                o = self._engine.depth_map(imgs[b], proj_matrices[b], depth_values[b], loss_gt=(depth_gt[b], mask[b]))
                valid = mask[b].to(torch.float32).sum() + 1e-6                      # drmvsnet.py:351
                losses.append(o["ce_sum"].to(torch.float32) / valid.to(o["ce_sum"].device))
                depths.append(o["depth"].clone())
        return torch.stack(losses).mean(), torch.stack(depths)


def mvsnet_cls_loss(prob_volume, depth_gt, mask, depth_value, return_prob_map=False):
    """Training loss of the reference (models/drmvsnet.py:347-381), kept as plain tensor ops as SURVEY.md §8 row a16
    prescribes: masked one-hot cross entropy of the probability volume against the hypothesis nearest to the ground
    truth, plus the winner-take-all depth map.  prob_volume [B,D,H,W], depth_gt / mask [B,H,W], depth_value [B,D]."""
    b, d, h, w = prob_volume.shape
    planes = depth_value.view(b, d, 1, 1)
    gt_index = (planes - depth_gt.unsqueeze(1)).abs().argmin(dim=1)
    gt_index = torch.round(mask * gt_index.to(prob_volume.dtype)).long().unsqueeze(1)        # masked pixels -> plane 0
    cross_entropy = -torch.log(prob_volume.gather(1, gt_index)).squeeze(1)
    valid = mask.sum(dim=[1, 2]) + 1e-6
    loss = ((mask * cross_entropy).sum(dim=[1, 2]) / valid).mean()
    wta_depth = planes.expand(b, d, h, w).gather(1, prob_volume.argmax(dim=1, keepdim=True)).squeeze(1)
    if return_prob_map:
        return loss, wta_depth, prob_volume.max(dim=1)[0]
    return loss, wta_depth
