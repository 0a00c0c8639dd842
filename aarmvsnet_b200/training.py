"""One optimisation step of the reference's training loop (train.py:288-307) on the B200 kernels:

    optimizer.zero_grad(); prob = model(imgs, proj_matrices, depth_values); loss = mvsnet_cls_loss(prob, depth_gt, mask,
    depth_values); loss.backward(); optimizer.step()

`FlatAdam` keeps all 90 parameters of an EMVSNet as views into ONE contiguous fp32 buffer (and their gradients likewise), so
the Adam update is a single kernel launch (csrc/adam.cu, torch.optim.Adam arithmetic); `TrainStep` chains the differentiable
forward (model.py: _RegularisedCostVolume), the loss (plain tensor ops, as SURVEY.md §8 row a16 prescribes) and that update.
The evidential loss of the fork (train.py:304) is out of scope: the loss is upstream's mvsnet_cls_loss.
"""
import ctypes

import torch
import torch.distributed as dist

from . import capi
from .engine import _ptr
from .model import mvsnet_cls_loss


class FlatAdam:
    """torch.optim.Adam(model.parameters(), lr) (train.py:180) over a flat parameter buffer."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = list(model.parameters())
        if not params or params[0].device.type != "cuda":
            raise RuntimeError("FlatAdam needs the model on a CUDA device (there is no CPU path)")
        dev = params[0].device
        n = sum(p.numel() for p in params)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        off = 0
        for p in params:                       # re-seat every parameter (and its .grad) as a view of the flat buffers
            k = p.numel()
            self.flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.grad[off:off + k].view_as(p)
            off += k
        self.model, self.params, self.n = model, params, n
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.steps = 0
        self.lib = capi.load()

    def zero_grad(self):
        self.grad.zero_()
        for p in self.params:                  # autograd accumulates into the views
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr() or p.grad.data_ptr() >= self.grad.data_ptr() + 4 * self.n:
                raise RuntimeError("a parameter's .grad was replaced: FlatAdam owns the gradient buffer")

    def step(self):
        self.steps += 1
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)
        with torch.cuda.device(self.flat.device):
            capi.check(self.lib.aarmvs_adam_step(_ptr(self.flat), _ptr(self.grad), _ptr(self.exp_avg), _ptr(self.exp_avg_sq), self.n,
                                                 self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.steps, stream),
                       "aarmvs_adam_step")
        # the kernel wrote through raw pointers: tensor version counters did not move, so tell the engine to re-pack its weights
        if hasattr(self.model, "_engine"):
            self.model._engine._wkey = None


def average_gradients(flat_grad, group=None):
    """Data-parallel training over the GPUs of one box (one process per GPU, each with its own batch): ONE all-reduce of the
    flat gradient buffer (187 203 floats = 0.75 MB; NCCL over NVLink, gloo in the CPU test) replaces the per-parameter
    reductions of the nn.DataParallel the reference leaves commented out (train.py:129-150,173).  No-op for a single process."""
    if dist.is_available() and dist.is_initialized():
        world = dist.get_world_size(group)
        if world > 1:
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
            flat_grad.div_(world)
    return flat_grad


class TrainStep:
    """train.py:288-307 for one batch: returns the loss (device scalar) and the winner-take-all depth map."""

    def __init__(self, model, lr=1e-3):
        self.model = model
        self.opt = FlatAdam(model, lr=lr)

    def __call__(self, imgs, proj_matrices, depth_values, depth_gt, mask):
        self.opt.zero_grad()
        prob, _, _ = self.model(imgs, proj_matrices, depth_values)
        loss, depth = mvsnet_cls_loss(prob, depth_gt, mask, depth_values)
        loss.backward()
        average_gradients(self.opt.grad)
        self.opt.step()
        return loss.detach(), depth
