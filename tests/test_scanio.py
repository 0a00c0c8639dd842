"""Formats either side of the path: cam / pair parsers, crop + scale, PFM round trip (CPU)."""
import os

import numpy as np
import pytest


def _write_scan(tmp, n_views=4, h=72, w=88):
    from PIL import Image
    rng = np.random.default_rng(0)
    os.makedirs(tmp / "images"); os.makedirs(tmp / "cams")
    for v in range(n_views):
        Image.fromarray(rng.integers(0, 255, (h, w, 3), dtype=np.uint8)).save(tmp / "images" / f"{v:08d}.png")
        E = np.eye(4, dtype=np.float32); E[0, 3] = -30.0 * v
        K = np.array([[150.0, 0, w / 2], [0, 150.0, h / 2], [0, 0, 1]], dtype=np.float32)
        with open(tmp / "cams" / f"{v:08d}_cam.txt", "w") as f:
            f.write("extrinsic\n" + "\n".join(" ".join(f"{x:.6f}" for x in row) for row in E) + "\n\n")
            f.write("intrinsic\n" + "\n".join(" ".join(f"{x:.6f}" for x in row) for row in K) + "\n\n")
            f.write("425.0 2.5\n")
    with open(tmp / "pair.txt", "w") as f:
        f.write(f"{n_views}\n")
        for v in range(n_views):
            others = [u for u in range(n_views) if u != v]
            f.write(f"{v}\n{len(others)} " + " ".join(f"{u} {100.0 - u:.1f}" for u in others) + "\n")
    return tmp


def _golden_scan(tmp, case):
    """Rebuilds the scan folder of one case of tests/golden/scanio_ref.npz (bytes stored by oracle/make_golden_scanio.py)."""
    from conftest import GOLD
    g = np.load(os.path.join(GOLD, "scanio_ref.npz"))
    prefix = f"{case}/file/"
    for key in g.files:
        if key.startswith(prefix):
            p = tmp / key[len(prefix):]
            os.makedirs(p.parent, exist_ok=True)
            p.write_bytes(g[key].tobytes())
    return g


@pytest.mark.parametrize("case,kw", [("a", dict(view_num=3, numdepth=16, interval_scale=0.8, inverse_depth=False, max_h=96, max_w=128)),
                                     ("b", dict(view_num=3, numdepth=16, interval_scale=1.06, inverse_depth=False, max_h=64, max_w=80)),
                                     ("c", dict(view_num=5, numdepth=12, interval_scale=1.0, inverse_depth=True, max_h=72, max_w=88))])
def test_loader_matches_reference_dataset(tmp_path, case, kw):
    """eval_scan.build_sample (scanio parsers, standardisation, adaptive scale, centre crop, projection matrices, depth
    hypotheses) against what the UNMODIFIED reference loader MVSDataset.__getitem__ returned for the same scan folder
    (datasets/data_eval_transform.py:89-174, datasets/preprocess.py:26-63): bit for bit."""
    import argparse
    import eval_scan
    from aarmvsnet_b200 import scanio
    g = _golden_scan(tmp_path, case)
    pairs = scanio.read_pair_file(tmp_path / "pair.txt")
    assert len(pairs) == int(g[f"{case}/len"])
    args = argparse.Namespace(**kw)
    for idx in (0, len(pairs) - 1):
        ref, srcs = pairs[idx]
        ids, imgs, projs, depth_values = eval_scan.build_sample(str(tmp_path), ref, srcs, args)
        assert ids[0] == ref and len(ids) == g[f"{case}/{idx}/proj_matrices"].shape[0]
        np.testing.assert_array_equal(projs, g[f"{case}/{idx}/proj_matrices"])
        np.testing.assert_array_equal(depth_values, g[f"{case}/{idx}/depth_values"])
        if f"{case}/{idx}/imgs" in g.files:
            np.testing.assert_array_equal(imgs, g[f"{case}/{idx}/imgs"])


def test_pfm_writer_matches_reference_bytes(tmp_path):
    """scanio.save_pfm against the bytes the reference's save_pfm wrote (datasets/data_io.py:48-74)."""
    from conftest import GOLD
    from aarmvsnet_b200 import scanio
    g = np.load(os.path.join(GOLD, "scanio_ref.npz"))
    for key in ("grey", "color"):
        p = tmp_path / (key + ".pfm")
        scanio.save_pfm(p, g[f"pfm/{key}/array"])
        assert p.read_bytes() == g[f"pfm/{key}/bytes"].tobytes()
        back, scale = scanio.read_pfm(p)
        assert scale == 1.0 and np.array_equal(back, g[f"pfm/{key}/array"])


def test_pair_and_cam_parsers(tmp_path):
    from aarmvsnet_b200 import scanio
    scan = _write_scan(tmp_path)
    pairs = scanio.read_pair_file(scan / "pair.txt")
    assert len(pairs) == 4 and pairs[0] == (0, [1, 2, 3]) and pairs[2] == (2, [0, 1, 3])
    K, E, dmin, dint = scanio.read_cam_file(scan / "cams" / "00000002_cam.txt", interval_scale=0.4)
    assert K.shape == (3, 3) and E.shape == (4, 4) and K.dtype == np.float32
    assert E[0, 3] == -60.0 and K[0, 0] == 150.0 and dmin == 425.0 and abs(dint - 1.0) < 1e-6
    P = scanio.projection(K, E)
    np.testing.assert_allclose(P[:3, :4], K @ E[:3, :4])
    assert (P[3] == [0, 0, 0, 1]).all()
    img, *_ = scanio.load_view(str(scan), 1)
    assert img.shape == (72, 88, 3) and abs(img.mean()) < 1e-4 and abs(img[..., 0].std() - 1) < 1e-3


def test_depth_hypotheses_and_fit():
    from aarmvsnet_b200 import scanio
    d = scanio.depth_hypotheses(425.0, 1.0, 512, inverse_depth=False)
    assert d.shape == (512,) and d[0] == 425.0 and d[-1] == 936.0 and d.dtype == np.float32
    inv = scanio.depth_hypotheses(425.0, 1.0, 8, inverse_depth=True)
    assert abs(inv[0] - 425.0) < 1e-3 and np.all(np.diff(inv) > 0)
    np.testing.assert_allclose(np.diff(1.0 / inv), np.diff(1.0 / inv)[0], rtol=1e-4)      # uniform in 1/depth
    assert scanio.fit_scale([(1200, 1600)], 600, 800) == 0.5
    with pytest.raises(ValueError):
        scanio.fit_scale([(500, 800)], 600, 800)
    img = np.zeros((74, 91, 3), np.float32)
    K = np.array([[100.0, 0, 45.5], [0, 100.0, 37.0], [0, 0, 1]], np.float32)
    out, k2 = scanio.crop_view(img, K, 64, 88)
    assert out.shape[:2] == (64, 88) and k2[0, 2] == K[0, 2] - 2 and k2[1, 2] == K[1, 2] - 5
    out, k2 = scanio.crop_view(img, K, 600, 800)                 # fits: rounded down to multiples of 8
    assert out.shape[:2] == (72, 88)


def test_pfm_round_trip(tmp_path):
    from aarmvsnet_b200 import scanio
    a = np.random.default_rng(1).random((13, 17), dtype=np.float32)
    scanio.save_pfm(tmp_path / "sub" / "a.pfm", a)
    b, scale = scanio.read_pfm(tmp_path / "sub" / "a.pfm")
    assert scale == 1.0 and np.array_equal(a, b)
    raw = open(tmp_path / "sub" / "a.pfm", "rb").read()
    assert raw.startswith(b"Pf\n17 13\n-1.000000\n")             # grey, width height, little endian
    np.testing.assert_array_equal(np.frombuffer(raw[-17 * 4:], "<f4"), a[0])      # rows are stored bottom-up
    c = np.random.default_rng(2).random((5, 4, 3), dtype=np.float32)
    scanio.save_pfm(tmp_path / "c.pfm", c)
    assert np.array_equal(scanio.read_pfm(tmp_path / "c.pfm")[0], c)
    with pytest.raises(ValueError):
        scanio.save_pfm(tmp_path / "d.pfm", a.astype(np.float64))


@pytest.mark.gpu
def test_eval_scan_driver(tmp_path):
    """The driver end to end on a synthetic scan, with and without the per-scan feature cache."""
    import subprocess
    import sys
    from conftest import ROOT
    from aarmvsnet_b200 import scanio
    scan = _write_scan(tmp_path / "scan0", n_views=4, h=72, w=88)
    outs = []
    for extra in ([], ["--cache_features"]):
        out = tmp_path / ("out" + str(len(extra)))
        cmd = [sys.executable, os.path.join(ROOT, "eval_scan.py"), "--scan", str(scan), "--out", str(out), "--numdepth", "16",
               "--view_num", "3", "--max_h", "64", "--max_w", "80"] + extra
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        depth, _ = scanio.read_pfm(out / "depth_est" / "00000001.pfm")
        conf, _ = scanio.read_pfm(out / "confidence" / "00000001.pfm")
        assert depth.shape == (64, 80) and conf.shape == (64, 80) and (conf > 0).all() and (conf <= 1).all()
        assert len(os.listdir(out / "depth_est")) == 4
        outs.append((depth, conf))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.allclose(outs[0][1], outs[1][1], atol=1e-6)
    # several scans in one job: units = (scan, view) pairs, outputs under OUT/<scan folder>/; the feature cache is per scan
    scan1 = _write_scan(tmp_path / "scan1", n_views=3, h=72, w=88)
    out = tmp_path / "out_multi"
    cmd = [sys.executable, os.path.join(ROOT, "eval_scan.py"), "--scan", str(scan), str(scan1), "--out", str(out), "--numdepth", "16",
           "--view_num", "3", "--max_h", "64", "--max_w", "80", "--cache_features"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(os.listdir(out / "scan0" / "depth_est")) == 4 and len(os.listdir(out / "scan1" / "depth_est")) == 3
    depth, _ = scanio.read_pfm(out / "scan0" / "depth_est" / "00000001.pfm")
    assert np.array_equal(depth, outs[0][0])


@pytest.mark.gpu
def test_eval_scan_driver_against_reference_loader_and_oracle(tmp_path):
    """The whole caller chain against the reference's: eval_scan.py on the scan folder of golden case "b" must write the depth /
    confidence maps that the oracle forward (pinned to the reference model) produces from the sample the UNMODIFIED reference
    loader returned for that folder (tests/golden/scanio_ref.npz)."""
    import subprocess
    import sys
    import torch
    from conftest import ROOT, load_weights
    from aarmvsnet_b200 import scanio
    from oracle import aarmvs_oracle as orc
    g = _golden_scan(tmp_path / "scan_b", "b")
    out = tmp_path / "out"
    cmd = [sys.executable, os.path.join(ROOT, "eval_scan.py"), "--scan", str(tmp_path / "scan_b"), "--out", str(out), "--numdepth", "16",
           "--view_num", "3", "--max_h", "64", "--max_w", "80", "--interval_scale", "1.06", "--cache_features"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    sd = load_weights("model_dtu_v2.ckpt")
    for idx, ref_view in ((0, 0), (3, 3)):
        imgs = torch.from_numpy(g["b/0/imgs"])[None] if idx == 0 else None
        if imgs is None:
            continue                                      # images are stored for the first sample only
        ref = orc.forward_test(sd, imgs, torch.from_numpy(g[f"b/{idx}/proj_matrices"])[None],
                               torch.from_numpy(g[f"b/{idx}/depth_values"])[None])
        depth, _ = scanio.read_pfm(out / "depth_est" / f"{ref_view:08d}.pfm")
        conf, _ = scanio.read_pfm(out / "confidence" / f"{ref_view:08d}.pfm")
        same = torch.from_numpy(depth.copy()) == ref["depth"][0]
        assert same.float().mean().item() >= 0.999
        np.testing.assert_allclose(conf[same.numpy()], ref["photometric_confidence"][0].numpy()[same.numpy()], atol=1e-3, rtol=1e-3)
