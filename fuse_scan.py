#!/usr/bin/env python
"""Depth maps -> filtered point cloud: the step after the hot path (reference: fusion.py:135-273), B200 edition.

    python fuse_scan.py --scan DIR --out DIR --ply cloud.ply [--photo_threshold 0.35] [--depth_dir depth_est] [--conf_dir confidence]

Reads what eval_scan.py wrote (DIR/out: depth_est/{:08d}.pfm, confidence/{:08d}.pfm) plus the scan's pair.txt, cams/ and
images/, runs the geometric + photometric consistency filter of every reference view as ONE kernel launch
(aarmvsnet_b200.fusion.filter_view -> aarmvs_fuse_view) and writes a binary PLY.  One process per GPU (launch with
torch.distributed.run for several): rank r filters a contiguous slice of the reference views (no communication on the
data path), rank 0 concatenates the parts.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def fit_image_and_cam(img, conf_shape):
    """fusion.py:149-166: scale the image so that one side matches the depth map, centre-crop the other; returns the
    image and (scale, index, flag) for read_camera_parameters (fusion.py:26-41)."""
    from PIL import Image
    ch, cw = conf_shape
    scale = float(ch) / img.shape[0]
    index = int((int(img.shape[1] * scale) - cw) / 2)
    index_p = (int(img.shape[1] * scale) - cw) - index
    flag = 0
    if cw / img.shape[1] > scale:
        scale = float(cw) / img.shape[1]
        index = int((int(img.shape[0] * scale) - ch) / 2)
        index_p = (int(img.shape[0] * scale) - ch) - index
        flag = 1
    size = (int(img.shape[1] * scale), int(img.shape[0] * scale))
    img = np.asarray(Image.fromarray((img * 255).astype(np.uint8)).resize(size, Image.BILINEAR), dtype=np.float32) / 255.0
    img = img[:, index:img.shape[1] - index_p] if flag == 0 else img[index:img.shape[0] - index_p]
    return img, scale, index, flag


def camera(scan_dir, view, scale, index, flag):
    from aarmvsnet_b200 import scanio
    intr, extr, _, _ = scanio.read_cam_file(os.path.join(scan_dir, "cams", f"{view:08d}_cam.txt"))
    intr = intr.copy()
    intr[:2, :] *= scale
    intr[0 if flag == 0 else 1, 2] -= index
    return intr, extr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scan", required=True)
    ap.add_argument("--out", required=True, help="folder holding depth_est/ and confidence/ (written by eval_scan.py)")
    ap.add_argument("--ply", required=True)
    ap.add_argument("--photo_threshold", type=float, default=0.35)
    ap.add_argument("--depth_dir", default="depth_est")
    ap.add_argument("--conf_dir", default="confidence")
    args = ap.parse_args()

    from PIL import Image
    from aarmvsnet_b200 import fusion, scan as scan_mod, scanio
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("fuse_scan.py needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")

    pairs = scanio.read_pair_file(os.path.join(args.scan, "pair.txt"))
    mine = scan_mod.shard_views(len(pairs), world, rank)
    depth = lambda v: scanio.read_pfm(os.path.join(args.out, args.depth_dir, f"{v:08d}.pfm"))[0]      # noqa: E731
    pts, cols = [], []
    t0 = time.time()
    for idx in mine:
        ref, srcs = pairs[idx]
        if not os.path.exists(os.path.join(args.out, args.depth_dir, f"{ref:08d}.pfm")):
            print("skip", ref)
            continue
        srcs = [s for s in srcs if os.path.exists(os.path.join(args.out, args.depth_dir, f"{s:08d}.pfm"))][:10]
        if not srcs:
            continue
        d_ref = np.ascontiguousarray(depth(ref), dtype=np.float32)
        conf = np.ascontiguousarray(scanio.read_pfm(os.path.join(args.out, args.conf_dir, f"{ref:08d}.pfm"))[0], dtype=np.float32)
        img_path = next(p for p in (os.path.join(args.scan, "images", f"{ref:08d}{e}") for e in (".jpg", ".png", ".jpeg")) if os.path.exists(p))
        img = np.array(Image.open(img_path), dtype=np.float32) / 255.0
        img, scale, index, flag = fit_image_and_cam(img, conf.shape)
        K_ref, E_ref = camera(args.scan, ref, scale, index, flag)
        cams = [camera(args.scan, s, scale, index, flag) for s in srcs]
        d_src = np.stack([np.ascontiguousarray(depth(s), dtype=np.float32) for s in srcs])
        out = fusion.filter_view(torch.from_numpy(d_ref).to(dev), torch.from_numpy(conf).to(dev), K_ref, E_ref,
                                 torch.from_numpy(d_src).to(dev), [c[0] for c in cams], [c[1] for c in cams], args.photo_threshold)
        xyz, (ys, xs) = fusion.world_points(out["depth_avg"], out["final_mask"], K_ref, E_ref)
        pts.append(xyz.cpu().numpy())
        cols.append((img[ys.cpu().numpy(), xs.cpu().numpy()] * 255).astype(np.uint8))
        print(f"rank {rank}: ref-view {ref:02d}: photo/geo/final mask {out['photo_mask'].float().mean():.4f}/"
              f"{out['geo_mask'].float().mean():.4f}/{out['final_mask'].float().mean():.4f}")
    torch.cuda.synchronize()
    pts = np.concatenate(pts) if pts else np.zeros((0, 3))
    cols = np.concatenate(cols) if cols else np.zeros((0, 3), np.uint8)
    if world > 1:
        import torch.distributed as dist
        gathered = [None] * world
        dist.all_gather_object(gathered, (pts, cols))
        pts, cols = np.concatenate([g[0] for g in gathered]), np.concatenate([g[1] for g in gathered])
    if rank == 0:
        fusion.write_ply(args.ply, pts, cols)
        print(f"{len(pts)} points -> {args.ply}  ({time.time() - t0:.2f} s for {len(mine)} of {len(pairs)} views on rank 0)")


if __name__ == "__main__":
    main()
