"""Depth-map fusion filter (fusion.py:71-133, 179-225): the numpy oracle against golden outputs of
the unmodified reference (oracle/make_golden_fusion.py), and the CUDA kernel (C ABI) against both."""
import os

import numpy as np
import pytest

from oracle import fusion_oracle as fo

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.mark.parametrize("name", ["fusion_small", "fusion_mid"])
def test_oracle_remap_matches_cv2(name):
    g = load(name)
    got = fo.remap_bilinear(g["depth"][1], g["remap_x"], g["remap_y"])
    np.testing.assert_array_equal(got, g["remap_out"])


@pytest.mark.parametrize("name", ["fusion_small", "fusion_mid"])
def test_oracle_matches_reference(name):
    g = load(name)
    K, E, D = g["K"], g["E"], g["depth"]
    masks, mask, d_rep, x_src, y_src, vis = fo.check_geometric_consistency(D[0], K[0], E[0], D[1], K[1], E[1])
    np.testing.assert_array_equal(x_src, g["v1_x_src"])
    np.testing.assert_array_equal(y_src, g["v1_y_src"])
    np.testing.assert_array_equal(np.stack(masks), g["v1_masks"])
    np.testing.assert_array_equal(vis, g["v1_vis"])
    np.testing.assert_array_equal(d_rep, g["v1_depth_rep"])
    out = fo.filter_view(D[0], g["conf"], K[0], E[0], list(D[1:]), list(K[1:]), list(E[1:]), 0.35)
    np.testing.assert_array_equal(out["geo_mask"], g["geo_mask"])
    np.testing.assert_array_equal(out["geo_sum"], g["geo_sum"])
    np.testing.assert_array_equal(out["depth_avg"], g["depth_avg"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["fusion_small", "fusion_mid"])
def test_kernel_matches_reference(name):
    import torch
    from aarmvsnet_b200 import fusion
    g = load(name)
    K, E, D = g["K"], g["E"], g["depth"]
    out = fusion.filter_view(torch.from_numpy(D[0]).cuda(), torch.from_numpy(g["conf"]).cuda(), K[0], E[0],
                             torch.from_numpy(D[1:]).cuda(), K[1:], E[1:], photo_threshold=0.35)
    geo = out["geo_mask"].cpu().numpy().astype(bool)
    # fp64 per-pixel arithmetic in a different summation order than numpy's matmul: identical except for values that sit
    # within an ulp of a threshold or of a float32 rounding boundary
    assert (geo != g["geo_mask"]).mean() <= 2e-4
    assert (out["geo_sum"].cpu().numpy() != g["geo_sum"]).mean() <= 5e-4
    same = out["geo_sum"].cpu().numpy() == g["geo_sum"]
    np.testing.assert_allclose(out["depth_avg"].cpu().numpy()[same], g["depth_avg"][same], rtol=2e-6, atol=0)
    photo = g["conf"] > 0.35
    np.testing.assert_array_equal(out["final_mask"].cpu().numpy().astype(bool), np.logical_and(photo, geo))


@pytest.mark.gpu
def test_kernel_full_size_properties():
    """cfg2-sized maps: a view fused with exact copies of itself under identity relative pose keeps every valid pixel."""
    import torch
    from aarmvsnet_b200 import fusion
    h, w = 1184, 1600
    gen = torch.Generator().manual_seed(0)
    d = (500 + 100 * torch.rand(h, w, generator=gen)).cuda()
    d[:10] = 0
    K = np.array([[2892.0, 0, 800], [0, 2892.0, 592], [0, 0, 1]], np.float32)
    E = np.eye(4, dtype=np.float32)
    out = fusion.filter_view(d, torch.ones(h, w).cuda(), K, E, torch.stack([d, d, d]), np.stack([K] * 3), np.stack([E] * 3), 0.35)
    geo = out["geo_mask"].bool()
    assert geo[10:].all() and not geo[:10].any()
    torch.testing.assert_close(out["depth_avg"][10:], d[10:], rtol=1e-6, atol=0)


@pytest.mark.gpu
def test_fuse_scan_driver(tmp_path):
    """fuse_scan.py end to end on a synthetic scan folder: PLY vertex count and coordinates against the numpy oracle."""
    import subprocess
    import sys
    from PIL import Image
    from aarmvsnet_b200 import scanio
    g = load("fusion_small")
    K, E, D, conf = g["K"], g["E"], g["depth"], g["conf"]
    n, (h, w) = len(D), D[0].shape
    scan, out = tmp_path / "scan", tmp_path / "out"
    for d in (scan / "images", scan / "cams", out / "depth_est", out / "confidence"):
        d.mkdir(parents=True)
    rng = np.random.default_rng(0)
    with open(scan / "pair.txt", "w") as f:
        f.write(f"{n}\n")
        for v in range(n):
            others = [u for u in range(n) if u != v]
            f.write(f"{v}\n{len(others)} " + " ".join(f"{u} 1.0" for u in others) + "\n")
    for v in range(n):
        Image.fromarray(rng.integers(0, 255, (h, w, 3), dtype=np.uint8)).save(scan / "images" / f"{v:08d}.png")
        with open(scan / "cams" / f"{v:08d}_cam.txt", "w") as f:
            f.write("extrinsic\n" + "\n".join(" ".join(repr(float(x)) for x in row) for row in E[v]) + "\n\nintrinsic\n")
            f.write("\n".join(" ".join(repr(float(x)) for x in row) for row in K[v]) + "\n\n425.0 2.5\n")
        scanio.save_pfm(str(out / "depth_est" / f"{v:08d}.pfm"), D[v])
        scanio.save_pfm(str(out / "confidence" / f"{v:08d}.pfm"), conf)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ply = tmp_path / "cloud.ply"
    subprocess.run([sys.executable, os.path.join(root, "fuse_scan.py"), "--scan", str(scan), "--out", str(out), "--ply", str(ply)],
                   check=True, timeout=300)
    raw = open(ply, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    count = int([ln for ln in head.decode().splitlines() if ln.startswith("element vertex")][0].split()[-1])
    rec = np.frombuffer(body, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1")])
    assert len(rec) == count
    want_pts = []
    for v in range(n):
        others = [u for u in range(n) if u != v]
        o = fo.filter_view(D[v], conf, K[v], E[v], [D[u] for u in others], [K[u] for u in others], [E[u] for u in others], 0.35)
        want_pts.append(fo.world_points(o["depth_avg"], o["final_mask"], K[v], E[v]))
    want = np.concatenate(want_pts)
    assert abs(count - len(want)) <= max(2, int(2e-4 * len(want)))
    if count == len(want):
        np.testing.assert_allclose(np.stack([rec["x"], rec["y"], rec["z"]], 1), want.astype(np.float32), rtol=1e-5, atol=1e-3)
