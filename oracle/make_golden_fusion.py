"""Generates tests/golden/fusion_*.npz from the UNMODIFIED reference fusion.py (imported from /root/reference with
stub modules for its absent plotting / PLY dependencies; cv2 is the real one).  Run in the build container only:

    python oracle/make_golden_fusion.py
"""
import os
import sys
import types

import statistics, random  # noqa: F401,E401  (stdlib modules the reference tree shadows: import them first)
import cv2  # noqa: F401
import numpy as np
import PIL.Image  # noqa: F401
import torch, torchvision  # noqa: F401,E401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(self.__name__ + "." + k)

    def __call__(self, *a, **k):
        return None


def load_reference():
    sys.path.insert(0, "/root/reference")
    argv, sys.argv = sys.argv, ["fusion.py"]
    for n in ("plyfile", "matplotlib", "matplotlib.pyplot", "tensorboardX", "torchviz", "torchsummary"):
        sys.modules.setdefault(n, _Stub(n))
    import fusion
    sys.argv = argv
    return fusion


def scene(h, w, nviews, seed, noise=0.004, holes=True):
    """Depth maps of a tilted plane seen from a ring of cameras (K, E float32 as read_camera_parameters returns them),
    with multiplicative noise around the thresholds of the filter and a few invalid (zero) regions."""
    rng = np.random.default_rng(seed)
    f = 1.8 * w
    Ks, Es, depths = [], [], []
    nrm, c = np.array([0.05, -0.03, 1.0]), 600.0                     # plane nrm . X = c
    for v in range(nviews):
        ang = 0.0 if v == 0 else rng.uniform(-0.12, 0.12)
        ang2 = 0.0 if v == 0 else rng.uniform(-0.05, 0.05)
        Ry = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        Rx = np.array([[1, 0, 0], [0, np.cos(ang2), -np.sin(ang2)], [0, np.sin(ang2), np.cos(ang2)]])
        R = Rx @ Ry
        centre = np.array([0.0, 0.0, 0.0]) if v == 0 else np.array([-600.0 * np.sin(ang) + rng.uniform(-20, 20), rng.uniform(-40, 40), 600.0 * (1 - np.cos(ang))])
        t = -R @ centre
        E = np.eye(4)
        E[:3, :3], E[:3, 3] = R, t
        K = np.array([[f, 0, w / 2 + rng.uniform(-3, 3)], [0, f, h / 2 + rng.uniform(-3, 3)], [0, 0, 1]])
        K32, E32 = K.astype(np.float32), E.astype(np.float32)
        ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
        rays = np.linalg.inv(K) @ np.stack([xs.ravel(), ys.ravel(), np.ones(h * w)])
        # X = R^T (d ray - t);  nrm . X = c  ->  d = (c + nrm . R^T t) / (nrm . R^T ray)
        d = (c + nrm @ (R.T @ t)) / (nrm @ (R.T @ rays))
        d = d.reshape(h, w) * (1.0 + noise * rng.standard_normal((h, w)))
        if holes:
            y0, x0 = rng.integers(0, h - 8), rng.integers(0, w - 8)
            d[y0:y0 + 6, x0:x0 + 7] = 0.0
        Ks.append(K32); Es.append(E32); depths.append(d.astype(np.float32))
    conf = rng.uniform(0, 1, (h, w)).astype(np.float32)
    return Ks, Es, depths, conf


def main():
    ref = load_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    for name, (h, w, nviews, seed, noise) in {"fusion_small": (48, 64, 4, 1, 0.004), "fusion_mid": (96, 128, 11, 2, 0.003)}.items():
        Ks, Es, depths, conf = scene(h, w, nviews, seed, noise)
        n = nviews                                                        # n = len(src_views) + 1 (fusion.py:198-200)
        rec = {"K": np.stack(Ks), "E": np.stack(Es), "depth": np.stack(depths), "conf": conf}
        geo_sum, sums, drs = 0, [], []
        for ct, v in enumerate(range(1, nviews), start=1):
            masks, geo, d_rep, x_src, y_src, vis = ref.check_geometric_consistency(depths[0], Ks[0], Es[0], depths[v], Ks[v], Es[v])
            if ct == 1:
                sums = [masks[i - 2].astype(np.int32) for i in range(2, n)]
                rec["v1_masks"] = np.stack(masks)
                rec["v1_depth_rep"], rec["v1_x_src"], rec["v1_y_src"], rec["v1_vis"] = d_rep, x_src, y_src, vis
            else:
                for i in range(2, n):
                    sums[i - 2] += masks[i - 2].astype(np.int32)
            geo_sum = geo_sum + geo.astype(np.int32)
            drs.append(d_rep)
        geo_mask = geo_sum >= n                                           # fusion.py:219-224
        for i in range(2, n):
            geo_mask = np.logical_or(geo_mask, sums[i - 2] >= i)
        rec["geo_mask"], rec["geo_sum"] = geo_mask, geo_sum
        rec["depth_avg"] = ((sum(drs) + depths[0]) / (geo_sum + 1)).astype(np.float32)
        # cv2.remap itself on rough coordinates, incl. far outside / NaN, to pin the restated interpolation
        rng = np.random.default_rng(seed + 100)
        mx = rng.uniform(-3, w + 3, (h, w)).astype(np.float32)
        my = rng.uniform(-3, h + 3, (h, w)).astype(np.float32)
        mx[0, :4] = [np.nan, 1e9, -1e9, w - 1]
        my[0, :4] = [0.0, 0.0, 1.0, h - 1]
        rec["remap_x"], rec["remap_y"] = mx, my
        rec["remap_out"] = cv2.remap(depths[1], mx, my, interpolation=cv2.INTER_LINEAR)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(name, "geo_mask mean", geo_mask.mean(), "geo_sum max", geo_sum.max())


if __name__ == "__main__":
    main()
