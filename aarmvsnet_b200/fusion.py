"""Depth-map fusion filter on the GPU: the per-reference-view part of the reference's
filter_depth (fusion.py:135-255) -- geometric consistency against every source view, photometric mask, averaged depth,
world points -- as one kernel launch per reference view (`aarmvs_fuse_view`, csrc/fusion.cu).  No CPU path.
"""
import ctypes

import numpy as np
import torch

from . import capi


def _view_mats(K_ref, E_ref, K_src, E_src):
    """The six matrices of fusion.py:80-105 with the reference's dtype flow (float32 inverses / products of the float32
    camera matrices), flattened to the kernel's [60] float64 record."""
    K_ref, E_ref, K_src, E_src = (np.asarray(m, np.float32) for m in (K_ref, E_ref, K_src, E_src))
    parts = [np.linalg.inv(K_ref), np.matmul(E_src, np.linalg.inv(E_ref))[:3], K_src,
             np.linalg.inv(K_src), np.matmul(E_ref, np.linalg.inv(E_src))[:3], K_ref]
    return np.concatenate([p.astype(np.float64).reshape(-1) for p in parts])


def filter_view(depth_ref, conf_ref, K_ref, E_ref, depth_src, K_src, E_src, photo_threshold):
    """depth_ref, conf_ref: [H,W] float32 CUDA tensors; depth_src: [S,H,W]; K_*: 3x3, E_*: 4x4 float32 (numpy, as
    read_camera_parameters returns them, fusion.py:26-41).  Returns CUDA tensors: geo_mask, photo_mask, final_mask (uint8),
    depth_avg (float32), geo_sum (int32)."""
    if depth_ref.device.type != "cuda":
        raise RuntimeError("aarmvsnet_b200 has no CPU path: tensors must live on a CUDA device")
    lib = capi.load()
    H, W = depth_ref.shape
    S = depth_src.shape[0]
    if not 1 <= S <= 10:
        raise RuntimeError("1..10 source views (the reference keeps nine nested consistency thresholds)")
    dev = depth_ref.device
    with torch.cuda.device(dev):
        mats = torch.from_numpy(np.stack([_view_mats(K_ref, E_ref, K_src[v], E_src[v]) for v in range(S)])).to(dev)
        assert mats.shape[1] == lib.aarmvs_fuse_view_mats()
        dr = depth_ref.to(torch.float32).contiguous()
        cf = conf_ref.to(torch.float32).contiguous()
        ds = depth_src.to(torch.float32).contiguous()
        geo = torch.empty(H, W, device=dev, dtype=torch.uint8)
        fin = torch.empty(H, W, device=dev, dtype=torch.uint8)
        avg = torch.empty(H, W, device=dev, dtype=torch.float32)
        gsum = torch.empty(H, W, device=dev, dtype=torch.int32)
        p = lambda t: ctypes.c_void_p(t.data_ptr())     # noqa: E731
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        capi.check(lib.aarmvs_fuse_view(p(dr), p(cf), p(ds), S, H, W, p(mats), float(photo_threshold), p(geo), p(fin), p(avg),
                                        p(gsum), stream), "aarmvs_fuse_view")
    return {"geo_mask": geo, "photo_mask": (cf > photo_threshold).to(torch.uint8), "final_mask": fin, "depth_avg": avg,
            "geo_sum": gsum}


def world_points(depth_avg, mask, K_ref, E_ref):
    """fusion.py:244-252: surviving pixels -> [M,3] float64 world coordinates (row-major pixel order), on the device."""
    dev = depth_avg.device
    ys, xs = torch.nonzero(mask.bool(), as_tuple=True)
    d = depth_avg[ys, xs].to(torch.float64)
    Kinv = torch.from_numpy(np.linalg.inv(np.asarray(K_ref, np.float32)).astype(np.float64)).to(dev)
    Einv = torch.from_numpy(np.linalg.inv(np.asarray(E_ref, np.float32)).astype(np.float64)).to(dev)
    pix = torch.stack([xs.to(torch.float64) * d, ys.to(torch.float64) * d, d])
    cam = Kinv @ pix
    world = Einv @ torch.cat([cam, torch.ones_like(d)[None]])
    return world[:3].t().contiguous(), (ys, xs)


def write_ply(path, points, colors):
    """Binary little-endian PLY with float x, y, z and uchar red, green, blue per vertex (what plyfile writes for the
    reference's vertex_all, fusion.py:255-268)."""
    pts = np.asarray(points, dtype="<f4")
    col = np.asarray(colors, dtype=np.uint8)
    rec = np.empty(len(pts), dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"), ("blue", "u1")])
    rec["x"], rec["y"], rec["z"] = pts[:, 0], pts[:, 1], pts[:, 2]
    rec["red"], rec["green"], rec["blue"] = col[:, 0], col[:, 1], col[:, 2]
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
              "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n" % len(pts))
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(rec.tobytes())
