"""CPU restatement (numpy) of the reference's depth-map fusion filter: fusion.py:71-133 (reproject_with_depth,
check_geometric_consistency) and the per-view accumulation of filter_depth (fusion.py:179-225, 244-255).

TEST INFRASTRUCTURE ONLY: imported by tests/ and oracle/make_golden_fusion.py, never by the product path.
Pinned: tests/golden/fusion_*.npz hold the outputs of the UNMODIFIED reference functions (imported from /root/reference by
oracle/make_golden_fusion.py, cv2.remap included); tests/test_fusion.py checks this restatement against them.

cv2.remap(INTER_LINEAR, BORDER_CONSTANT) is restated from OpenCV's remapBilinear for float images: map coordinates are
rounded to 1/32 pixel (cvRound(x * 32), round-half-even), the four weights come from the float table
(1 - a/32) and (a/32), and out-of-image taps contribute the border value 0.
"""
import numpy as np

INTER_BITS, INTER_TAB = 5, 32


def remap_bilinear(img, map_x, map_y):
    """cv2.remap(img, map_x, map_y, cv2.INTER_LINEAR) for a float32 single-channel image (border constant 0)."""
    h, w = img.shape
    with np.errstate(invalid="ignore", over="ignore"):
        fx = np.rint(map_x.astype(np.float32) * np.float32(INTER_TAB))
        fy = np.rint(map_y.astype(np.float32) * np.float32(INTER_TAB))
    # cvRound of NaN / out-of-range values is INT_MIN; such samples fall outside the image either way
    bad = ~np.isfinite(fx) | ~np.isfinite(fy) | (np.abs(fx) > 2.0e9) | (np.abs(fy) > 2.0e9)
    sx = np.where(bad, -(2 ** 31), fx).astype(np.int64)
    sy = np.where(bad, -(2 ** 31), fy).astype(np.int64)
    ix = np.clip(sx >> INTER_BITS, -32768, 32767)
    iy = np.clip(sy >> INTER_BITS, -32768, 32767)
    ax = (sx & (INTER_TAB - 1)).astype(np.float32) * np.float32(1.0 / INTER_TAB)
    ay = (sy & (INTER_TAB - 1)).astype(np.float32) * np.float32(1.0 / INTER_TAB)
    wx0, wx1 = np.float32(1.0) - ax, ax
    wy0, wy1 = np.float32(1.0) - ay, ay

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        return np.where(ok, img[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], np.float32(0))

    v00, v01, v10, v11 = tap(iy, ix), tap(iy, ix + 1), tap(iy + 1, ix), tap(iy + 1, ix + 1)
    out = v00 * (wy0 * wx0) + v01 * (wy0 * wx1) + v10 * (wy1 * wx0) + v11 * (wy1 * wx1)
    return out.astype(np.float32)


def view_matrices(K_ref, E_ref, K_src, E_src):
    """The six matrices of fusion.py:80-105 with numpy's dtype flow: inverses and products of the float32 camera matrices
    are float32, everything they multiply is float64."""
    K_ref, E_ref, K_src, E_src = (np.asarray(m, np.float32) for m in (K_ref, E_ref, K_src, E_src))
    return {
        "Kr_inv": np.linalg.inv(K_ref), "T_rs": np.matmul(E_src, np.linalg.inv(E_ref)), "Ks": K_src,
        "Ks_inv": np.linalg.inv(K_src), "T_sr": np.matmul(E_ref, np.linalg.inv(E_src)), "Kr": K_ref,
    }


def reproject_with_depth(depth_ref, K_ref, E_ref, depth_src, K_src, E_src):
    """fusion.py:71-112."""
    h, w = depth_ref.shape
    m = view_matrices(K_ref, E_ref, K_src, E_src)
    x_ref, y_ref = np.meshgrid(np.arange(0, w), np.arange(0, h))
    x_ref, y_ref = x_ref.reshape(-1), y_ref.reshape(-1)
    one = np.ones_like(x_ref)
    xyz_ref = np.matmul(m["Kr_inv"], np.vstack((x_ref, y_ref, one)) * depth_ref.reshape(-1))
    xyz_src = np.matmul(m["T_rs"], np.vstack((xyz_ref, one)))[:3]
    k_xyz_src = np.matmul(m["Ks"], xyz_src)
    with np.errstate(divide="ignore", invalid="ignore"):
        xy_src = k_xyz_src[:2] / k_xyz_src[2:3]
    x_src = xy_src[0].reshape(h, w).astype(np.float32)
    y_src = xy_src[1].reshape(h, w).astype(np.float32)
    sampled = remap_bilinear(depth_src, x_src, y_src)
    xyz_src = np.matmul(m["Ks_inv"], np.vstack((xy_src, one)) * sampled.reshape(-1))
    xyz_rep = np.matmul(m["T_sr"], np.vstack((xyz_src, one)))[:3]
    depth_rep = xyz_rep[2].reshape(h, w).astype(np.float32)
    k_xyz_rep = np.matmul(m["Kr"], xyz_rep)
    with np.errstate(divide="ignore", invalid="ignore"):
        xy_rep = k_xyz_rep[:2] / k_xyz_rep[2:3]
    return depth_rep, xy_rep[0].reshape(h, w).astype(np.float32), xy_rep[1].reshape(h, w).astype(np.float32), x_src, y_src


def check_geometric_consistency(depth_ref, K_ref, E_ref, depth_src, K_src, E_src):
    """fusion.py:114-133: nine nested (distance, relative depth) thresholds i/4 px and i/1300, i = 2..10; the last one is
    the view's geometric mask; the reprojected depth is zeroed outside it."""
    h, w = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, w), np.arange(0, h))
    depth_rep, x_rep, y_rep, x_src, y_src = reproject_with_depth(depth_ref, K_ref, E_ref, depth_src, K_src, E_src)
    with np.errstate(divide="ignore", invalid="ignore"):
        dist = np.sqrt((x_rep - x_ref) ** 2 + (y_rep - y_ref) ** 2)             # float64 (float32 - int64)
        rel = np.abs(depth_rep - depth_ref) / depth_ref                          # float32
        masks = [np.logical_and(dist < i / 4, rel < np.float32(i / 1300)) for i in range(2, 11)]
        vis = np.logical_and(dist < 1, rel < np.float32(0.01))
    mask = masks[-1]
    depth_rep = depth_rep.copy()
    depth_rep[~mask] = 0
    return masks, mask, depth_rep, x_src, y_src, vis


def filter_view(depth_ref, conf_ref, K_ref, E_ref, src_depths, src_Ks, src_Es, photo_threshold):
    """The per-reference-view part of filter_depth (fusion.py:179-225): accumulate the nested masks over the source views,
    geo_mask = OR_i (count_i >= i) for i = 2..n-1 (n = views + 1), average the consistent depths, AND with the photometric mask."""
    n = len(src_depths) + 1
    geo_sum = np.zeros(depth_ref.shape, np.int32)
    sums = [np.zeros(depth_ref.shape, np.int32) for _ in range(2, n)]
    acc = np.zeros(depth_ref.shape, np.float32)
    for d_s, k_s, e_s in zip(src_depths, src_Ks, src_Es):
        masks, geo, d_rep, _, _, _ = check_geometric_consistency(depth_ref, K_ref, E_ref, d_s, k_s, e_s)
        for i in range(2, n):
            sums[i - 2] += masks[i - 2].astype(np.int32)
        geo_sum += geo.astype(np.int32)
        acc = acc + d_rep
    geo_mask = geo_sum >= n
    for i in range(2, n):
        geo_mask = np.logical_or(geo_mask, sums[i - 2] >= i)
    depth_avg = (acc + depth_ref) / (geo_sum + 1)
    photo = conf_ref > photo_threshold
    return {"geo_mask": geo_mask, "photo_mask": photo, "final_mask": np.logical_and(photo, geo_mask),
            "depth_avg": depth_avg.astype(np.float32), "geo_sum": geo_sum}


def world_points(depth_avg, mask, K_ref, E_ref):
    """fusion.py:244-252: back-project the surviving pixels to world coordinates -> [M, 3] float64."""
    h, w = depth_avg.shape
    x, y = np.meshgrid(np.arange(0, w), np.arange(0, h))
    x, y, d = x[mask], y[mask], depth_avg[mask]
    K_ref, E_ref = np.asarray(K_ref, np.float32), np.asarray(E_ref, np.float32)
    xyz_ref = np.matmul(np.linalg.inv(K_ref), np.vstack((x, y, np.ones_like(x))) * d)
    xyz_world = np.matmul(np.linalg.inv(E_ref), np.vstack((xyz_ref, np.ones_like(x))))[:3]
    return xyz_world.transpose((1, 0))
