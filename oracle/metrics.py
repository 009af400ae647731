"""Oracle: repeatability metrics and the homography helpers they use.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  NumPy restatement (vectorised, float64 like the reference's
Python floats) of
  balf/benchmark_test/repeatability_tools.py:379-490   compute_repeatability (+ intersection_area :492-510,
                                                       union_area :512-513)
  balf/benchmark_test/repeatability_tools.py:516-614   compute_resize_repeatability
  balf/benchmark_test/geometry_tools.py:7-27           create_common_region_masks
  balf/benchmark_test/geometry_tools.py:43-64          apply_homography_to_points (+ getAff :66-84)
PINNED: ``oracle/make_golden.py r2`` ran the reference's own functions on seeded inputs
(tests/golden/r2_metrics.npz); tests/test_oracle_golden.py re-checks this file against them.  The reference sorts with
NumPy's unstable argsort; as everywhere in this repo the canonical order is the stable one (ties keep index order) and
the golden vectors were produced with ``np.argsort`` forced stable.
``cv2.warpPerspective`` (OpenCV, a pip dependency of the reference that IS installed in the image) is restated in
``warp_perspective_linear`` following OpenCV's imgwarp.cpp (INTER_LINEAR with 5 fractional bits, constant border) and
checked against ``cv2`` itself by the CPU tests.
"""
import numpy as np

_EPS32 = float(np.finfo(np.float32).eps)
_EPS64 = float(np.finfo(float).eps)


# ------------------------------------------------------------------------------------------ circle overlap
def intersection_area(R, r, d):
    """repeatability_tools.py:492-510, element-wise on arrays (R, r, d broadcastable)."""
    R, r, d = np.broadcast_arrays(np.asarray(R, np.float64), np.asarray(r, np.float64), np.asarray(d, np.float64))
    out = np.zeros(d.shape, np.float64)
    inside = d <= np.abs(R - r)
    apart = (~inside) & (d >= r + R)
    mid = ~(inside | apart)
    out[inside] = np.pi * np.minimum(R, r)[inside] ** 2
    if mid.any():
        Rm, rm, dm = R[mid], r[mid], d[mid]
        r2, R2, d2 = rm ** 2, Rm ** 2, dm ** 2
        alpha = np.arccos((d2 + r2 - R2) / (2 * dm * rm))
        beta = np.arccos((d2 + R2 - r2) / (2 * dm * Rm))
        out[mid] = r2 * alpha + R2 * beta - 0.5 * (r2 * np.sin(2 * alpha) + R2 * np.sin(2 * beta))
    return out


def union_area(r, R, inter):
    return (np.pi * (r ** 2)) + (np.pi * (R ** 2)) - inter


def overlap_matrices(src, dst, eps=1e-6, dist_match_thresh=3, radious_size=30.0):
    """The double loop of compute_repeatability (:398-427): (multi-scale overlaps, single-scale overlaps,
    possible_matches)."""
    src = np.asarray(src, np.float64)
    dst = np.asarray(dst, np.float64)
    dx = src[:, None, 0] - dst[None, :, 0]
    dy = src[:, None, 1] - dst[None, :, 1]
    dist = ((dx ** 2) + (dy ** 2)) ** 0.5
    possible = int((dist <= dist_match_thresh).any(axis=1).sum()) if dst.shape[0] else 0
    near = dist <= 4 * radious_size
    rr = np.broadcast_to(src[:, None, 2], dist.shape)
    rd = np.broadcast_to(dst[None, :, 2], dist.shape)
    factor = radious_size / (np.maximum(rr, rd) + _EPS64)
    inter = intersection_area(factor * rr, factor * rd, dist)
    multi = np.where(near, inter / (union_area(factor * rr, factor * rd, inter) + eps), 0.0)
    rs = np.full(dist.shape, float(radious_size))
    inter = intersection_area(rs, rs, dist)
    single = np.where(near, inter / (union_area(rs, rs, inter) + eps), 0.0)
    return multi, single, possible


def greedy_assign(m, min_overlap):
    """:432-447 -- walk the entries in descending order (stable), keep a pair when its row and column are free,
    stop at the first free entry below ``min_overlap``.  -> (count, summed (1 - overlap), [[x_pos, y_pos], ...])."""
    n_src, n_dst = m.shape
    order = np.argsort((-1 * m).flatten(), kind="stable")
    yv = np.zeros(n_src, bool)
    xv = np.zeros(n_dst, bool)
    found, err, corr = 0, 0.0, []
    for index in order:
        y, x = divmod(int(index), n_dst)
        if xv[x] or yv[y]:
            continue
        v = m[y, x]
        if v < min_overlap:
            break
        found += 1
        err += (1 - v)
        corr.append([x, y])
        xv[x] = yv[y] = True
    return found, err, np.asarray(corr)


def compute_repeatability(src_indexes, dst_indexes, overlap_err=0.4, eps=1e-6, dist_match_thresh=3, radious_size=30.):
    src = np.asarray(src_indexes, np.float64)
    dst = np.asarray(dst_indexes, np.float64)
    multi, single, possible = overlap_matrices(src, dst, eps, dist_match_thresh, radious_size)
    fs, es, cs = greedy_assign(single, 1 - overlap_err)
    fm, em, cm = greedy_assign(multi, 1 - overlap_err)
    points = min(len(src), len(dst))
    with np.errstate(divide="ignore", invalid="ignore"):
        rep_s = (fs / np.asarray(points, float)) * 100.0
        rep_m = (fm / np.asarray(points, float)) * 100.0
    em = 0.0 if fm == 0 else em / float(fm + _EPS64)
    es = 0.0 if fs == 0 else es / float(fs + _EPS64)
    return {'rep_single_scale': rep_s, 'rep_multi_scale': rep_m, 'num_points_single_scale': fs,
            'num_points_multi_scale': fm, 'error_overlap_single_scale': es, 'error_overlap_multi_scale': em,
            'total_num_points': points, 'correspondences': cs, 'possible_matches': possible, 'correspondences_m': cm}


# ------------------------------------------------------------------------------------------ SuperPoint-style repeatability
def _warp(points, H):
    hom = np.concatenate([points, np.ones((points.shape[0], 1))], axis=1)
    w = np.dot(hom, np.transpose(H))
    return w[:, :2] / w[:, 2:]


def _select_k_best(points, k):
    srt = points[np.argsort(points[:, 2], kind="stable"), :2]
    return srt[-min(k, points.shape[0]):, :] if points.shape[0] else srt


def compute_resize_repeatability(keypoints, warped_keypoints, h, shape_src, shape_dst, keep_k_points=1000,
                                 distance_thresh=5):
    """:516-614.  keypoints / warped_keypoints: [N,3] (row, col, prob).  The reference overwrites the first two columns
    of ``keypoints`` in place (:560-561); this restatement works on copies."""
    keypoints = np.array(keypoints, np.float64)
    warped_keypoints = np.array(warped_keypoints, np.float64)
    H = np.asarray(h, np.float64)
    wp = _warp(warped_keypoints[:, [1, 0]], np.linalg.inv(H))[:, [1, 0]]
    keep = (wp[:, 0] >= 0) & (wp[:, 0] < shape_src[0]) & (wp[:, 1] >= 0) & (wp[:, 1] < shape_src[1])
    warped_keypoints = warped_keypoints[keep, :]
    tw = _warp(keypoints[:, [1, 0]], H)
    true_warped = np.stack([tw[:, 1], tw[:, 0], keypoints[:, 2]], axis=-1)
    keep = (true_warped[:, 0] >= 0) & (true_warped[:, 0] < shape_dst[0]) & \
           (true_warped[:, 1] >= 0) & (true_warped[:, 1] < shape_dst[1])
    true_warped = true_warped[keep, :]
    warped_keypoints = _select_k_best(warped_keypoints, keep_k_points)
    true_warped = _select_k_best(true_warped, keep_k_points)
    n1, n2 = true_warped.shape[0], warped_keypoints.shape[0]
    norm = np.linalg.norm(true_warped[:, None, :] - warped_keypoints[None, :, :], ord=None, axis=2)
    count1 = count2 = 0
    le1 = le2 = None
    if n2 != 0:
        min1 = np.min(norm, axis=1)
        count1 = int(np.sum(min1 <= distance_thresh))
        le1 = min1[min1 <= distance_thresh]
    if n1 != 0:
        min2 = np.min(norm, axis=0)
        count2 = int(np.sum(min2 <= distance_thresh))
        le2 = min2[min2 <= distance_thresh]
    repeatability = (count1 + count2) / (n1 + n2) * 100.0 if n1 + n2 > 0 else 0
    localization_err = -1
    if count1 + count2 > 0:
        localization_err = 0
        if le1 is not None:
            localization_err += le1.sum() / (count1 + count2)
        if le2 is not None:
            localization_err += le2.sum() / (count1 + count2)
    else:
        repeatability = 0.
    return {'repeatability': repeatability, 'localization_err': localization_err, 'common_src_num': n1,
            'common_dst_num': n2, 'rep_src_num': count1, 'rep_dst_num': count2}


# ------------------------------------------------------------------------------------------ homography helpers
def apply_homography_to_points(points, h):
    """geometry_tools.py:43-64: position through h, radius through the local affine approximation
    (new radius = 1 / (e0 e1)^(1/4), e = eig((A M^-1 A^T)^-1), M^-1 = (r^2 + eps32) I)."""
    pts = np.asarray(points, np.float64).reshape(-1, 4)
    h = np.asarray(h, np.float64)
    if pts.shape[0] == 0:
        return np.asarray([])
    x, y = pts[:, 0], pts[:, 1]
    new = (h @ np.stack([x, y, np.ones_like(x)], 0)).T
    tmp = pts[:, 2] ** 2 + _EPS32
    den = h[2, 0] * x + h[2, 1] * y + h[2, 2]
    nx = h[0, 0] * x + h[0, 1] * y + h[0, 2]
    ny = h[1, 0] * x + h[1, 1] * y + h[1, 2]
    aff = np.empty((len(x), 2, 2))
    aff[:, 0, 0] = h[0, 0] / den - nx * h[2, 0] / den ** 2
    aff[:, 0, 1] = h[0, 1] / den - nx * h[2, 1] / den ** 2
    aff[:, 1, 0] = h[1, 0] / den - ny * h[2, 0] / den ** 2
    aff[:, 1, 1] = h[1, 1] / den - ny * h[2, 1] / den ** 2
    mi1_inv = np.linalg.inv(np.eye(2)[None] * (1 / tmp)[:, None, None])
    bmb = np.linalg.inv(aff @ (mi1_inv @ np.transpose(aff, (0, 2, 1))))
    e = np.linalg.eigvals(bmb)
    rad = 1 / ((e[:, 0] * e[:, 1]) ** 0.5) ** 0.5
    return np.stack([new[:, 0] / new[:, 2], new[:, 1] / new[:, 2], np.real(rad), pts[:, 3]], 1)


def remove_borders(image, borders):
    out = np.zeros_like(image)
    out[borders:image.shape[0] - borders, borders:image.shape[1] - borders] = \
        image[borders:image.shape[0] - borders, borders:image.shape[1] - borders]
    return out


def warp_perspective_linear(src, M, dsize):
    """cv2.warpPerspective(src, M, (w, h)) for a float64 single-channel image: INTER_LINEAR, BORDER_CONSTANT (0).
    Follows OpenCV imgwarp.cpp (WarpPerspectiveInvoker + remapBilinear): the inverse map is evaluated in double per
    64 x 16 block as (M0*bx + M1*y + M2) + M0*x1, scaled by INTER_TAB_SIZE / W, rounded to nearest-even to 1/32 pixel;
    the four taps are weighted with the float table (1 - a/32, a/32)."""
    src = np.asarray(src, np.float64)
    w, h = dsize
    Mi = np.linalg.inv(np.asarray(M, np.float64))
    m = Mi.ravel()
    xs = np.arange(w)
    bx = (xs // 64) * 64 if w >= 64 else np.zeros_like(xs)
    x1 = (xs - bx).astype(np.float64)
    bx = bx.astype(np.float64)
    ys = np.arange(h, dtype=np.float64)[:, None]
    X0 = (m[0] * bx[None, :] + m[1] * ys + m[2]) + m[0] * x1[None, :]
    Y0 = (m[3] * bx[None, :] + m[4] * ys + m[5]) + m[3] * x1[None, :]
    W0 = (m[6] * bx[None, :] + m[7] * ys + m[8]) + m[6] * x1[None, :]
    with np.errstate(divide="ignore"):
        Wi = np.where(W0 != 0, 32.0 / W0, 0.0)
    lim = float(2 ** 31)
    X = np.rint(np.clip(X0 * Wi, -lim, lim - 1)).astype(np.int64)
    Y = np.rint(np.clip(Y0 * Wi, -lim, lim - 1)).astype(np.int64)
    sx, ax = X >> 5, (X & 31).astype(np.float32)
    sy, ay = Y >> 5, (Y & 31).astype(np.float32)
    one = np.float32(1.0)
    wx1, wy1 = ax / np.float32(32), ay / np.float32(32)
    wx0, wy0 = one - wx1, one - wy1
    hs, ws = src.shape

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < hs) & (xx >= 0) & (xx < ws)
        return np.where(ok, src[np.clip(yy, 0, hs - 1), np.clip(xx, 0, ws - 1)], 0.0)
    return (tap(sy, sx) * (wx0 * wy0).astype(np.float64) + tap(sy, sx + 1) * (wx1 * wy0).astype(np.float64) +
            tap(sy + 1, sx) * (wx0 * wy1).astype(np.float64) + tap(sy + 1, sx + 1) * (wx1 * wy1).astype(np.float64))


def create_common_region_masks(h_dst_2_src, shape_src, shape_dst, borders=15):
    """geometry_tools.py:7-27."""
    h = np.asarray(h_dst_2_src, np.float64)
    inv_h = np.linalg.inv(h)
    inv_h = inv_h / inv_h[2, 2]
    ones_dst = remove_borders(np.ones((shape_dst[0], shape_dst[1])), borders)
    mask_src = warp_perspective_linear(ones_dst, h, (shape_src[1], shape_src[0]))
    mask_src = remove_borders(np.where(mask_src >= 0.75, 1.0, 0.0), borders)
    ones_src = remove_borders(np.ones((shape_src[0], shape_src[1])), borders)
    mask_dst = warp_perspective_linear(ones_src, inv_h, (shape_dst[1], shape_dst[0]))
    mask_dst = remove_borders(np.where(mask_dst >= 0.75, 1.0, 0.0), borders)
    return mask_src, mask_dst
