"""Generates tests/golden/scanio_ref.npz from the UNMODIFIED reference loader and PFM writer
(datasets/data_eval_transform.py MVSDataset.__getitem__, datasets/preprocess.py, datasets/data_io.py save_pfm), imported
from /root/reference with stub modules for its absent plotting dependency.  Run in the build container only:

    python oracle/make_golden_scanio.py

The fixture holds the BYTES of a small synthetic scan folder (so the test rebuilds exactly the same files; the images are
PNG streams stored under the `.jpg` names the reference hard-codes -- PIL decodes by content, and PNG is lossless, so no
JPEG-decoder version can change the pixels), the loader arguments of each case and what the reference returned:
`imgs` (first sample of a case), `proj_matrices`, `depth_values` per sample; plus the bytes save_pfm wrote for two arrays.
"""
import io
import os
import sys
import tempfile
import types

import statistics, random  # noqa: F401,E401  (stdlib modules the reference tree shadows: import them first)
import numpy as np
import PIL.Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(self.__name__ + "." + k)

    def __call__(self, *a, **k):
        return None


def load_reference():
    sys.path.insert(0, "/root/reference")          # before site-packages: a `datasets` package is installed there too
    for n in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(n, _Stub(n))
    from datasets import data_eval_transform, data_io
    return data_eval_transform.MVSDataset, data_io


def scan_files(n_views, h, w, seed):
    """{relative path: bytes} of a synthetic scan folder: smooth random images, a ring of cameras, a pair file."""
    rng = np.random.default_rng(seed)
    files = {}
    for v in range(n_views):
        img = rng.integers(0, 255, (h // 4 + 1, w // 4 + 1, 3), dtype=np.uint8)
        img = np.asarray(PIL.Image.fromarray(img).resize((w, h), PIL.Image.BILINEAR))
        buf = io.BytesIO()
        PIL.Image.fromarray(img).save(buf, format="PNG")
        files[f"images/{v:08d}.jpg"] = buf.getvalue()
        ang = 0.07 * v
        E = np.eye(4)
        E[:3, :3] = [[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]
        E[:3, 3] = [-35.0 * v, 4.0 * v, 2.0 * v]
        K = np.array([[1.7 * w, 0, w / 2 + 0.37 * v], [0, 1.7 * w, h / 2 - 0.21 * v], [0, 0, 1]])
        txt = "extrinsic\n" + "\n".join(" ".join(f"{x:.6f}" for x in row) for row in E) + "\n\n"
        txt += "intrinsic\n" + "\n".join(" ".join(f"{x:.6f}" for x in row) for row in K) + "\n\n"
        txt += "425.0 2.5\n"
        files[f"cams/{v:08d}_cam.txt"] = txt.encode()
    pair = f"{n_views}\n"
    for v in range(n_views):
        others = sorted((u for u in range(n_views) if u != v), key=lambda u: abs(u - v))
        pair += f"{v}\n{len(others)} " + " ".join(f"{u} {100.0 - abs(u - v):.1f}" for u in others) + "\n"
    files["pair.txt"] = pair.encode()
    return files


def write_scan(root, name, files):
    for rel, data in files.items():
        p = os.path.join(root, name, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "wb") as f:
            f.write(data)


# name: (views in the folder, image h, w, loader kwargs).  Sizes chosen so that the adaptive scale is < 1 and the crop cuts
# columns (a), nothing is scaled and the size is already a multiple of 8 (b), and inverse depths with fewer source views
# than requested (c).
CASES = {
    "a": (5, 150, 210, dict(nviews=3, ndepths=16, interval_scale=0.8, inverse_depth=False, max_h=96, max_w=128)),
    "b": (4, 64, 80, dict(nviews=3, ndepths=16, interval_scale=1.06, inverse_depth=False, max_h=64, max_w=80)),
    "c": (3, 80, 96, dict(nviews=5, ndepths=12, interval_scale=1.0, inverse_depth=True, max_h=72, max_w=88)),
}


def main():
    MVSDataset, data_io = load_reference()
    rec = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (n_views, h, w, kw) in CASES.items():
            files = scan_files(n_views, h, w, seed=ord(name))
            write_scan(tmp, "scan_" + name, files)
            with open(os.path.join(tmp, f"list_{name}.txt"), "w") as f:
                f.write("scan_" + name + "\n")
            for rel, data in files.items():
                rec[f"{name}/file/{rel}"] = np.frombuffer(data, dtype=np.uint8)
            ds = MVSDataset(tmp, os.path.join(tmp, f"list_{name}.txt"), "test", **kw)
            rec[f"{name}/len"] = np.int64(len(ds))
            for idx in (0, len(ds) - 1):
                s = ds[idx]
                if idx == 0:                                   # the images of the first sample only (fixture size)
                    rec[f"{name}/{idx}/imgs"] = s["imgs"]
                rec[f"{name}/{idx}/proj_matrices"] = s["proj_matrices"]
                rec[f"{name}/{idx}/depth_values"] = s["depth_values"]
                print(name, idx, s["imgs"].shape, s["proj_matrices"].shape, s["depth_values"][:3], s["filename"])
        rng = np.random.default_rng(5)
        grey = rng.random((13, 17), dtype=np.float32) * 900.0
        color = rng.random((6, 5, 3), dtype=np.float32)
        for key, arr in (("grey", grey), ("color", color)):
            p = os.path.join(tmp, key + ".pfm")
            data_io.save_pfm(p, arr)
            rec[f"pfm/{key}/array"] = arr
            rec[f"pfm/{key}/bytes"] = np.frombuffer(open(p, "rb").read(), dtype=np.uint8)
            back, scale = data_io.read_pfm(p)
            assert np.array_equal(back, arr) and scale == 1.0
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "scanio_ref.npz"), **rec)
    print("scanio_ref.npz", os.path.getsize(os.path.join(ROOT, "tests", "golden", "scanio_ref.npz")), "bytes")


if __name__ == "__main__":
    main()
