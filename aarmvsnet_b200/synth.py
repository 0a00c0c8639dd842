"""Seeded synthetic MVS samples (images, cameras, depth hypotheses).

One generator shared by the CUDA path, the oracle, the tests and bench.py.
Shapes follow the reference's sample dict (datasets/data_eval_transform.py:171-174):
``imgs [B,N,3,H,W]``, ``proj_matrices [B,N,4,4]`` (rows 0-2 = K @ [R|t], row 3 = 0 0 0 1, view 0 =
reference; datasets/dtu_yao.py:144-145) and ``depth_values [B,D]``.
"""
import math

import torch

DEPTH_MIN = 425.0
DEPTH_MAX = 935.0


def depth_hypotheses(num_depth, mode="uniform", depth_min=DEPTH_MIN, depth_max=DEPTH_MAX):
    """``uniform``: equally spaced planes (eval_dtu.sh: 425 + k * interval).
    ``inverse``: equally spaced in 1/depth (datasets/dtu_yao.py:153-157)."""
    if mode == "uniform":
        d = torch.linspace(depth_min, depth_max, num_depth, dtype=torch.float64)
    elif mode == "inverse":
        d = 1.0 / torch.linspace(1.0 / depth_min, 1.0 / depth_max, num_depth, dtype=torch.float64)
    else:
        raise ValueError(mode)
    return d.to(torch.float32)


def _rot_yx(ay, ax):
    cy, sy, cx, sx = math.cos(ay), math.sin(ay), math.cos(ax), math.sin(ax)
    ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=torch.float64)
    rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=torch.float64)
    return ry @ rx


def cameras(num_views, height, width, gen, depth_min=DEPTH_MIN, depth_max=DEPTH_MAX):
    """Ring of cameras looking at the point (0, 0, z_mid) of the reference frustum."""
    f = 1.8 * width
    K = torch.tensor([[f, 0, width / 2.0], [0, f, height / 2.0], [0, 0, 1]], dtype=torch.float64)
    z_mid = 0.5 * (depth_min + depth_max)
    projs = []
    for v in range(num_views):
        E = torch.eye(4, dtype=torch.float64)
        if v > 0:
            u = torch.rand(3, generator=gen, dtype=torch.float64)
            ay = (0.08 + 0.17 * u[0].item()) * (1.0 if u[1].item() < 0.5 else -1.0)
            ax = (u[2].item() - 0.5) * 0.1
            R = _rot_yx(ay, ax)  # world -> camera
            target = torch.tensor([0.0, 0.0, z_mid], dtype=torch.float64)
            centre = target - z_mid * (R.T @ torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64))
            E[:3, :3] = R
            E[:3, 3] = -R @ centre
        P = E.clone()
        P[:3, :4] = K @ E[:3, :4]
        projs.append(P)
    return torch.stack(projs).to(torch.float32)


def make_sample(height, width, num_views, num_depth, seed=0, batch=1, depth_mode="uniform",
                smooth=False):
    """Returns dict(imgs, proj_matrices, depth_values) of CPU fp32 tensors."""
    gen = torch.Generator().manual_seed(seed)
    imgs = torch.randn(batch, num_views, 3, height, width, generator=gen, dtype=torch.float32)
    if smooth:
        # low-pass variant: decisive cost minima instead of white noise
        k = torch.ones(1, 1, 5, 5) / 25.0
        flat = imgs.reshape(-1, 1, height, width)
        flat = torch.nn.functional.conv2d(flat, k, padding=2)
        flat = (flat - flat.mean(dim=(2, 3), keepdim=True)) / (flat.std(dim=(2, 3), keepdim=True) + 1e-8)
        imgs = flat.reshape(batch, num_views, 3, height, width).contiguous()
    projs = torch.stack([cameras(num_views, height, width, gen) for _ in range(batch)])
    depth_values = depth_hypotheses(num_depth, depth_mode)[None].repeat(batch, 1).contiguous()
    return {"imgs": imgs, "proj_matrices": projs, "depth_values": depth_values}
