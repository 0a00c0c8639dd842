#!/usr/bin/env python
"""Scan -> depth maps: the caller of the hot path (reference: eval.py:56-169), B200 edition.

    python eval_scan.py --scan DIR [DIR ...] --out DIR --ckpt tests/golden/model_dtu_v2.ckpt --numdepth 512 --view_num 7 \\
                        --max_h 1184 --max_w 1600 [--interval_scale 0.4] [--inverse_depth] [--cache_features]

One process per GPU (launch with torch.distributed.run for several): rank r handles a contiguous slice of the
(scan, reference view) pairs of all scans (aarmvsnet_b200.scan.shard_units) and writes depth_est/{:08d}.pfm +
confidence/{:08d}.pfm for them (under OUT/<scan folder name>/ when several scans are given); there is no communication.  Timing is CUDA-synchronised (the reference's print is not, eval.py:112-120).
--cache_features runs FeatNet once per image of the scan instead of once per (reference view, image).
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def build_sample(scan_dir, ref, srcs, args):
    from aarmvsnet_b200 import scanio
    ids = [ref] + srcs[: args.view_num - 1]
    views = [scanio.load_view(scan_dir, v, args.interval_scale) for v in ids]
    scale = fit_scale = scanio.fit_scale([v[0].shape[:2] for v in views], args.max_h, args.max_w)
    imgs, projs = [], []
    for img, intr, extr, _, _ in views:
        img, intr = scanio.scale_view(img, intr, scale)
        img, intr = scanio.crop_view(img, intr, args.max_h, args.max_w)
        imgs.append(img.transpose(2, 0, 1))
        projs.append(scanio.projection(intr, extr))
    depth_values = scanio.depth_hypotheses(views[0][3], views[0][4], args.numdepth, args.inverse_depth)
    return ids, np.stack(imgs).astype(np.float32), np.stack(projs).astype(np.float32), depth_values


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scan", required=True, nargs="+", help="one or more scan folders (images/, cams/, pair.txt)")
    ap.add_argument("--out", required=True)
    ap.add_argument("--ckpt", default=os.path.join(ROOT, "tests", "golden", "model_dtu_v2.ckpt"))
    ap.add_argument("--numdepth", type=int, default=512)
    ap.add_argument("--view_num", type=int, default=7)
    ap.add_argument("--max_h", type=int, default=1184)
    ap.add_argument("--max_w", type=int, default=1600)
    ap.add_argument("--interval_scale", type=float, default=1.0)
    ap.add_argument("--inverse_depth", action="store_true")
    ap.add_argument("--cache_features", action="store_true")
    args = ap.parse_args()

    import aarmvsnet_b200 as aa
    from aarmvsnet_b200 import scan as scan_mod, scanio
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    scans = [os.path.normpath(p) for p in args.scan]
    pairs = [scanio.read_pair_file(os.path.join(p, "pair.txt")) for p in scans]
    out_dirs = [args.out if len(scans) == 1 else os.path.join(args.out, os.path.basename(p)) for p in scans]
    mine = scan_mod.shard_units([len(p) for p in pairs], world, rank)
    model = None
    feature_cache, cache_scan = {}, -1
    t_all = time.perf_counter()
    for si, i in mine:
        if si != cache_scan:                 # image ids repeat across scans
            feature_cache, cache_scan = {}, si
        ref, srcs = pairs[si][i]
        ids, imgs, projs, depth_values = build_sample(scans[si], ref, srcs, args)
        if model is None:
            model = aa.EMVSNet(32, image_scale=1.0, max_h=imgs.shape[2], max_w=imgs.shape[3], return_depth=True)
            sd = torch.load(args.ckpt, map_location="cpu", weights_only=False)["model"]
            model.load_state_dict({k[7:] if k.startswith("module.") else k: v for k, v in sd.items()}, strict=True)
            model = model.cuda().eval()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        feats = None
        if args.cache_features:
            feats = []
            for vid, img in zip(ids, imgs):
                if vid not in feature_cache:
                    feature_cache[vid] = model.features(torch.from_numpy(img).cuda())
                feats.append(feature_cache[vid])
        out = model._engine.depth_map(torch.from_numpy(imgs).cuda(), torch.from_numpy(projs), torch.from_numpy(depth_values).cuda(),
                                      features=feats)
        depth, conf = out["depth"].cpu().numpy(), out["photometric_confidence"].cpu().numpy()      # D2H = the sync point
        dt = time.perf_counter() - t0
        scanio.save_pfm(os.path.join(out_dirs[si], "depth_est", f"{ref:08d}.pfm"), depth)
        scanio.save_pfm(os.path.join(out_dirs[si], "confidence", f"{ref:08d}.pfm"), conf)
        print(f"[rank {rank}] {os.path.basename(scans[si])} view {ref}: {imgs.shape[3]}x{imgs.shape[2]} N={len(ids)} D={len(depth_values)} forward {dt * 1e3:.1f} ms")
    torch.cuda.synchronize()
    print(f"[rank {rank}] {len(mine)} depth maps in {time.perf_counter() - t_all:.2f} s")


if __name__ == "__main__":
    main()
