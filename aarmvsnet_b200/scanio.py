"""On-disk formats either side of the hot path: MVSNet-style scan folders in, PFM maps out.

Follows the conventions of the reference's eval loader (datasets/data_eval_transform.py:34-69,89-174,
datasets/preprocess.py:7-73) and PFM I/O (datasets/data_io.py:9-74); host-side numpy only.
    <scan>/pair.txt                     view selection
    <scan>/cams/{:08d}_cam.txt          extrinsic 4x4, intrinsic 3x3, depth_min depth_interval
    <scan>/images/{:08d}.jpg|png
"""
import math
import os
import re
import sys

import numpy as np


def read_pair_file(path):
    """-> [(ref_view, [src_view, ...]), ...] best source first (data_eval_transform.py:40-50)."""
    with open(path) as f:
        tokens = f.read().split("\n")
    lines = [ln.strip() for ln in tokens if ln.strip() != ""]
    n = int(lines[0])
    out = []
    for i in range(n):
        ref = int(lines[1 + 2 * i])
        fields = lines[2 + 2 * i].split()
        out.append((ref, [int(x) for x in fields[1::2]]))           # "k id score id score ..."
    return out


def read_cam_file(path, interval_scale=1.0):
    """-> (intrinsics [3,3] f32, extrinsics [4,4] f32, depth_min, depth_interval * interval_scale)
    (data_eval_transform.py:55-69: extrinsic on lines 1-4, intrinsic on lines 7-9, depth range on line 11)."""
    with open(path) as f:
        lines = [ln.rstrip() for ln in f.readlines()]
    extr = np.array(" ".join(lines[1:5]).split(), dtype=np.float32).reshape(4, 4)
    intr = np.array(" ".join(lines[7:10]).split(), dtype=np.float32).reshape(3, 3)
    rng = lines[11].split()
    return intr, extr, float(rng[0]), float(rng[1]) * interval_scale


def standardize_image(img):
    """Per-channel zero mean / unit variance (data_eval_transform.py:79-83)."""
    img = np.asarray(img, dtype=np.float32)
    return (img - img.mean(axis=(0, 1), keepdims=True)) / np.sqrt(img.var(axis=(0, 1), keepdims=True))


def depth_hypotheses(depth_min, depth_interval, num_depth, inverse_depth):
    """data_eval_transform.py:119-129: uniform planes, or uniform in 1/depth from 1/depth_min towards 0 (exclusive)."""
    if inverse_depth:
        inv = np.linspace(1.0 / depth_min, 0.0, num_depth, endpoint=False)
        return (1.0 / inv).astype(np.float32)
    return np.arange(depth_min, depth_interval * num_depth + depth_min, depth_interval, dtype=np.float32)[:num_depth]


def fit_scale(shapes, max_h, max_w):
    """Adaptive down-scale factor so every view fits max_h x max_w (data_eval_transform.py:135-151)."""
    h_scale = max(float(max_h) / h for h, _ in shapes)
    w_scale = max(float(max_w) / w for _, w in shapes)
    if h_scale > 1 or w_scale > 1:
        raise ValueError("max_h, max_w must not exceed the image size")
    return max(h_scale, w_scale)


def scale_view(img, intr, scale):
    """Bilinear resize + intrinsics scaling (preprocess.py:7-24)."""
    import cv2
    if scale != 1:
        img = cv2.resize(img, None, fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
    k = intr.copy()
    k[0, 0] *= scale; k[1, 1] *= scale; k[0, 2] *= scale; k[1, 2] *= scale
    return img, k


def crop_view(img, intr, max_h, max_w, base=8):
    """Centre crop to at most max_h x max_w and a multiple of `base`; shifts the principal point (preprocess.py:41-63).
    (The reference rounds the size UP to a multiple of `base`, which only works if it already is one; here the size is
    rounded DOWN so that odd sizes crop instead of failing.)"""
    h, w = img.shape[:2]
    new_h = max_h if h > max_h else (h // base) * base
    new_w = max_w if w > max_w else (w // base) * base
    sh, sw = int(math.ceil((h - new_h) / 2)), int(math.ceil((w - new_w) / 2))
    k = intr.copy()
    k[0, 2] -= sw; k[1, 2] -= sh
    return img[sh:sh + new_h, sw:sw + new_w], k


def projection(intr, extr):
    """proj = E with rows 0-2 replaced by K @ E[:3,:4] (data_eval_transform.py:163-167, dtu_yao.py:144-145)."""
    p = extr.copy()
    p[:3, :4] = intr @ extr[:3, :4]
    return p


def read_pfm(path):
    """-> (array [H,W] or [H,W,3] float32, scale).  PFM stores rows bottom-up (data_io.py:9-45)."""
    with open(path, "rb") as f:
        header = f.readline().decode("utf-8").rstrip()
        if header not in ("PF", "Pf"):
            raise ValueError("not a PFM file")
        m = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("utf-8"))
        if not m:
            raise ValueError("malformed PFM header")
        w, h = int(m.group(1)), int(m.group(2))
        scale = float(f.readline().rstrip())
        endian = "<" if scale < 0 else ">"
        data = np.fromfile(f, endian + "f")
    shape = (h, w, 3) if header == "PF" else (h, w)
    return np.flipud(data.reshape(shape)).astype(np.float32), abs(scale)


def save_pfm(path, image, scale=1):
    """data_io.py:48-74."""
    image = np.asarray(image)
    if image.dtype != np.float32:
        raise ValueError("image dtype must be float32")
    if image.ndim == 3 and image.shape[2] == 3:
        color = True
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        color = False
    else:
        raise ValueError("image must be H x W x 3, H x W x 1 or H x W")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    little = image.dtype.byteorder == "<" or (image.dtype.byteorder == "=" and sys.byteorder == "little")
    with open(path, "wb") as f:
        f.write(b"PF\n" if color else b"Pf\n")
        f.write(f"{image.shape[1]} {image.shape[0]}\n".encode())
        f.write(("%f\n" % (-scale if little else scale)).encode())
        np.flipud(image).tofile(f)


def load_view(scan_dir, view_id, interval_scale=1.0):
    """-> (standardised image [H,W,3] f32, intrinsics, extrinsics, depth_min, depth_interval)."""
    from PIL import Image
    img_path = None
    for ext in (".jpg", ".png", ".jpeg"):
        cand = os.path.join(scan_dir, "images", f"{view_id:08d}{ext}")
        if os.path.exists(cand):
            img_path = cand
            break
    if img_path is None:
        raise FileNotFoundError(f"no image for view {view_id} under {scan_dir}/images")
    img = standardize_image(np.array(Image.open(img_path), dtype=np.float32))
    intr, extr, dmin, dint = read_cam_file(os.path.join(scan_dir, "cams", f"{view_id:08d}_cam.txt"), interval_scale)
    return img, intr, extr, dmin, dint
