"""Oracle: NumPy restatement of BALF's score-map post-processing.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Parity: PINNED against
``/root/reference/balf/utils/test_utils.py`` (``tests/golden/postproc_*.npz``),
except ``soft_argmax_points`` whose last step calls torchgeometry (un-vendored,
restated in ``oracle/thirdparty.py`` -- PARITY UNPINNED for that step).

Tie rule.  All four sorts on the reference path are NumPy's default *unstable*
argsort (test_utils.py:111,134,165; demo_match.py:54), so the reference's order
among exactly-equal scores is implementation-defined.  This oracle (and the CUDA
path) use the canonical rule "score descending, then raster index y*W+x
ascending", i.e. a stable argsort of -score over raster-ordered candidates.
``oracle/make_golden.py`` runs the reference with ``np.argsort`` forced stable so
that both sides are comparable.
"""
import numpy as np

from . import thirdparty


# ----------------------------------------------------------------------------- D0 / P1 / P2
def make_shape_even(image):
    """test_utils.py:16-21 -- zero-pad bottom/right to even H, W.  image: [H,W,C]."""
    h, w = image.shape[:2]
    return np.pad(image, ((0, h & 1), (0, w & 1), (0, 0)))


def mod_pad_amounts(h, w, factor=64):
    """test_utils.py:23-28 -- total rows/cols added; zero when already a multiple."""
    ph = ((h + factor) // factor) * factor - h if h % factor else 0
    pw = ((w + factor) // factor) * factor - w if w % factor else 0
    return ph, pw


def mod_padding_symmetric(image, factor=64):
    """test_utils.py:23-32 -- ``pad//2`` zeros on each side (h, w are even here)."""
    ph, pw = mod_pad_amounts(image.shape[0], image.shape[1], factor)
    return np.pad(image, ((ph // 2, ph // 2), (pw // 2, pw // 2), (0, 0)))


def padded_geometry(h, w, factor=64):
    """Shapes used by demo_match.py:22-43.

    Returns (Hp, Wp, top, left, h_start, w_start): padded size, where the image
    sits inside the padded tensor, and where the un-pad crop starts.  The two
    coincide (top == h_start) -- kept separate because the reference computes
    them by different formulas.
    """
    he, we = h + (h & 1), w + (w & 1)
    ph, pw = mod_pad_amounts(he, we, factor)
    hp, wp = he + 2 * (ph // 2), we + 2 * (pw // 2)
    return hp, wp, ph // 2, pw // 2, hp // 2 - he // 2, wp // 2 - we // 2


def preprocess(im_u8, factor=64):
    """demo_match.py:22-29 -- /255. in float64, pad, cast fp32, HWC -> 1CHW (numpy)."""
    im = im_u8 / 255.0
    pad = mod_padding_symmetric(make_shape_even(im), factor)
    return np.ascontiguousarray(pad.astype(np.float32).transpose(2, 0, 1))[None]


def remove_borders(score, b):
    """test_utils.py:34-47 -- new array, zero outside [b:H-b, b:W-b]."""
    out = np.zeros_like(score)
    out[b:score.shape[0] - b, b:score.shape[1] - b] = score[b:score.shape[0] - b, b:score.shape[1] - b]
    return out


# ----------------------------------------------------------------------------- P7 / P8
def window_max(score, size):
    """scipy.ndimage.maximum_filter(footprint=ones((size,size)), mode='reflect') as used at
    test_utils.py:52.  Output i sees inputs [i - size//2, i + (size-1)//2]; 'reflect' only
    duplicates samples already inside that window, so it equals the window clipped to the map."""
    lo, hi = size // 2, (size - 1) // 2
    h, w = score.shape
    pad = np.full((h + lo + hi, w + lo + hi), -np.inf, dtype=score.dtype)
    pad[lo:lo + h, lo:lo + w] = score
    rows = pad[:, 0:w].copy()
    for d in range(1, lo + hi + 1):
        np.maximum(rows, pad[:, d:d + w], out=rows)
    out = rows[0:h].copy()
    for d in range(1, lo + hi + 1):
        np.maximum(out, rows[d:d + h], out=out)
    return out


def apply_nms(score, size):
    """test_utils.py:50-54 -- keep pixels equal to their window max (plateaus all survive)."""
    return score * (score == window_max(score, size))


def kth_value_threshold(score, k):
    """test_utils.py:76-88 -- k-th largest value of the whole map, with the <=0 fallbacks.
    Raises IndexError (like the reference) when the map has fewer than k elements."""
    desc = np.sort(score.ravel())[::-1]
    t = desc[k - 1]
    if t <= 0.0:
        pos = desc[desc > 0.0]
        t = pos[-1] if pos.size else score.dtype.type(0.0)
    return t


def find_index_higher_scores(score, num_points=1000, threshold=-1):
    """test_utils.py:74-95 -- (y, x) of the first ``num_points`` raster-order pixels >= t."""
    if threshold == -1:
        threshold = kth_value_threshold(score, num_points)
    ys, xs = np.nonzero(score >= threshold)
    return np.stack([ys, xs], 1)[:num_points]


def get_point_coordinates(score, scale_value=1.0, num_points=1000, threshold=-1, order_coord="xysr"):
    """test_utils.py:56-72 -- rows (x, y, scale, score) float64, raster order."""
    idx = find_index_higher_scores(score, num_points, threshold)
    ys, xs = idx[:, 0], idx[:, 1]
    a, b = (xs, ys) if order_coord == "xysr" else (ys, xs)
    out = np.empty((len(idx), 4), np.float64)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = a, b, scale_value, score[ys, xs]
    return out


def windowed_detect(score, border=15, nms_size=15, num_points=2048):
    """train_utils.py:446-452 -- border mask, windowed NMS, k-th-value top-k, sort by score.
    Returns [<=k, 4] (x, y, 1.0, score) under the canonical tie rule."""
    pts = get_point_coordinates(apply_nms(remove_borders(score, border), nms_size), num_points=num_points)
    order = np.argsort(-pts[:, 3], kind="stable")
    return pts[order][:num_points]


# ----------------------------------------------------------------------------- P3 / P4
def greedy_nms(xs, ys, scores, h, w, radius):
    """test_utils.py:130-168 (``nms_fast``) on integer candidates.

    Walk candidates by (score desc, input order asc); a candidate survives iff no earlier
    survivor lies within Chebyshev distance <= radius.  Returns indices into the input,
    in survivor order (= score-descending, canonical ties).  Python loop: small cases and
    the faithful CPU-baseline timing only; ``oracle/postproc_c.c`` is the fast checker."""
    n = len(xs)
    if n == 0:
        return np.zeros(0, np.int64)
    order = np.argsort(-np.asarray(scores, np.float64), kind="stable")
    if n == 1:
        return order
    alive = np.zeros((h + 2 * radius, w + 2 * radius), np.bool_)
    px = np.asarray(xs, np.int64)[order] + radius
    py = np.asarray(ys, np.int64)[order] + radius
    alive[py, px] = True
    keep = []
    for i in range(n):
        if alive[py[i], px[i]]:
            alive[py[i] - radius:py[i] + radius + 1, px[i] - radius:px[i] + radius + 1] = False
            keep.append(i)
    return order[np.asarray(keep, np.int64)]


def get_points_direct_from_score_map(heatmap, conf_thresh=0.015, nms_size=15, subpixel=True,
                                     patch_size=5, scale_value=1.0, order_coord="xysr", nms=greedy_nms):
    """test_utils.py:97-128 -- threshold (fp32 compare), greedy NMS, optional sub-pixel, pack."""
    h, w = heatmap.shape
    ys, xs = np.nonzero(heatmap >= conf_thresh)
    if len(xs) == 0:
        return np.zeros((0, 4))
    sc = heatmap[ys, xs].astype(np.float64)
    keep = nms(xs, ys, sc, h, w, nms_size)
    pts = np.stack([xs[keep].astype(np.float64), ys[keep].astype(np.float64), sc[keep]], 0)
    if subpixel:
        pts = soft_argmax_points(pts, heatmap, patch_size)
    a, b = (pts[0], pts[1]) if order_coord == "xysr" else (pts[1], pts[0])
    return np.stack([a, b, np.full_like(a, scale_value), pts[2]], 1)


# ----------------------------------------------------------------------------- P5
def extract_patches(heatmap, pts_xy_int, ps):
    """test_utils.py:184-196 -- ps x ps windows of the map zero-padded by int(ps/2); rows
    y - ps//2 ... y - ps//2 + ps - 1 (asymmetric for even ps)."""
    pad = int(ps / 2)
    hp = np.pad(heatmap, pad)
    return np.stack([hp[y:y + ps, x:x + ps] for x, y in pts_xy_int]) if len(pts_xy_int) else \
        np.zeros((0, ps, ps), heatmap.dtype)


def soft_argmax_points(pts, heatmap, patch_size=5):
    """test_utils.py:170-182 + 204-215 -- normalise patch by its sum (+1e-6), negatives -> 1e-6,
    log, then spatial soft-argmax (torchgeometry, restated -- unpinned); xy += centroid - ps//2."""
    pts = pts.copy()
    pat = extract_patches(heatmap, pts[:2].T.astype(int), patch_size).astype(np.float32)
    flat = pat.reshape(len(pat), -1)
    q = flat / (flat.sum(-1, keepdims=True, dtype=np.float32) + np.float32(1e-6))
    q[q < 0] = np.float32(1e-6)
    with np.errstate(divide="ignore"):
        lq = np.log(q)
    dxdy = thirdparty.spatial_soft_argmax2d(lq.reshape(-1, patch_size, patch_size))
    pts[:2] += dxdy.T.astype(np.float64) - patch_size // 2
    return pts


# ----------------------------------------------------------------------------- P6 (demo_match.detect tail)
def select_top(pts, num_features):
    """demo_match.py:51-57 -- sort rows by score desc (canonical ties), keep k, drop the score."""
    if pts.size == 0:
        return np.zeros((0, 3))
    return pts[np.argsort(-pts[:, 3], kind="stable")][:num_features, 0:3]


def box_nms(prob, size=4, iou=0.1, min_prob=0.015, keep_top_k=-1):
    """balf/benchmark_test/repeatability_tools.py:227-255 (box_nms) on one map [H,W] -> [H,W].

    torchvision.ops.nms restated: boxes (y - size/2, x - size/2, y + size/2, x + size/2) in float32 around every pixel with
    prob >= min_prob, visited in descending score order (ties: raster order -- the canonical rule), a box is dropped when its
    float32 IoU ``inter / (area_a + area_b - inter)`` with an already kept box exceeds ``iou``.  PINNED against
    ``torchvision.ops.nms`` (installed) by tests/test_oracle_golden.py and tests/golden/r2_boxnms.npz."""
    prob = np.asarray(prob, np.float32)
    h, w = prob.shape
    ys, xs = np.where(prob >= np.float32(min_prob))
    sc = prob[ys, xs]
    order = np.argsort(-sc, kind="stable")
    half = np.float32(size / 2.0)
    y1, x1 = ys.astype(np.float32) - half, xs.astype(np.float32) - half
    y2, x2 = ys.astype(np.float32) + half, xs.astype(np.float32) + half
    area = (y2 - y1) * (x2 - x1)
    reach = int(np.ceil(size))
    kept_grid = -np.ones((h, w), np.int64)
    keep = []
    thr = np.float32(iou)
    for i in order:
        y, x = int(ys[i]), int(xs[i])
        win = kept_grid[max(y - reach, 0):y + reach + 1, max(x - reach, 0):x + reach + 1]
        ok = True
        for j in win[win >= 0]:
            ih = max(np.float32(0), min(y2[i], y2[j]) - max(y1[i], y1[j]))
            iw = max(np.float32(0), min(x2[i], x2[j]) - max(x1[i], x1[j]))
            inter = np.float32(ih * iw)
            if np.float32(inter / np.float32(area[i] + area[j] - inter)) > thr:
                ok = False
                break
        if ok:
            kept_grid[y, x] = i
            keep.append(i)
    keep = np.asarray(keep, np.int64)
    if keep_top_k > 0 and len(keep) > keep_top_k:
        keep = keep[:keep_top_k]                      # already in descending score order
    out = np.zeros_like(prob)
    out[ys[keep], xs[keep]] = sc[keep]
    return out
