"""Drop-in for the reference's ``balf/utils/test_utils.py`` (NumPy in / NumPy out).

Same function names, argument meaning, return layouts (float64 ``[N,4] = (x, y, scale, score)``
point arrays, inputs never mutated) and error behaviour as test_utils.py:7-215; the arithmetic
runs in the CUDA kernels of ``balf_b200/csrc/nms.cu`` through the C-ABI.  Ordering among exactly
equal scores follows the canonical rule (score desc, raster index asc) -- the reference's own
order there is whatever NumPy's unstable argsort happens to give.

Pure data-layout helpers (zero padding, border masking of a host array, YAML) stay on the host:
they move bytes of a NumPy array and have nothing to accelerate; the fused device versions are
``balf_preprocess_u8`` and the ``border`` argument of the NMS entry points.
"""
import numpy as np
import torch
import yaml

from .. import _capi


def _dev():
    if not torch.cuda.is_available():
        raise RuntimeError("balf_b200 has no CPU path: a CUDA device is required")
    return torch.device("cuda", torch.cuda.current_device())


def _up(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(_dev())


def get_cfg_from_yaml_file(cfg_file):
    with open(cfg_file, "r") as f:
        return yaml.load(f, Loader=yaml.FullLoader)


# ------------------------------------------------------------------------------- layout helpers (host)
def make_shape_even(image):
    """test_utils.py:16-21"""
    h, w = image.shape[0], image.shape[1]
    return np.pad(image, ((0, h % 2), (0, w % 2), (0, 0)), mode="constant", constant_values=0)


def mod_padding_symmetric(image, factor=64):
    """test_utils.py:23-32, the reference's own arithmetic: pad to the NEXT multiple of ``factor`` unless the size already
    is one, ``pad // 2`` on each side -- so an odd size (the pipeline always calls make_shape_even first) gets an odd
    padded size, e.g. 121 rows -> 127 with (3, 3).  The device path (balf_pad_geometry) folds make_shape_even in."""
    h, w = image.shape[0], image.shape[1]
    padh = ((h + factor) // factor) * factor - h if h % factor != 0 else 0
    padw = ((w + factor) // factor) * factor - w if w % factor != 0 else 0
    return np.pad(image, ((padh // 2, padh // 2), (padw // 2, padw // 2), (0, 0)), mode="constant", constant_values=0)


def remove_borders(image, borders):
    """test_utils.py:34-47 -- returns a new array"""
    out = np.zeros_like(image)
    sl = (slice(borders, image.shape[0] - borders), slice(borders, image.shape[1] - borders))
    out[sl] = image[sl]
    return out


# ------------------------------------------------------------------------------- windowed path
def apply_nms(score_map, size):
    """test_utils.py:50-54 -- score * (score == size x size window maximum)"""
    out = _capi.apply_nms_map(_up(score_map), size).cpu().numpy()
    if score_map.dtype == np.float32:
        return out
    # other dtypes (float64 maps): the device decides which pixels equal their window maximum (in fp32, the precision of
    # every score map of the pipeline), the values returned are the caller's own
    return score_map * (out != 0)


def find_index_higher_scores(map, num_points=1000, threshold=-1):
    """test_utils.py:74-95 -- (y, x) rows, raster order, at most num_points."""
    if threshold != -1:
        return np.argwhere(map >= threshold)[:num_points]          # explicit threshold: plain host filter
    if map.size < num_points:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (num_points - 1, map.size))
    if num_points > K_MAX:
        raise ValueError("num_points = %d exceeds the %d points one call of the device top-k returns (the reference has no "
                         "limit)" % (num_points, K_MAX))
    xy, _, cnt = _capi.windowed_nms_topk(_up(map), num_points, border=0, nms_size=1)
    n = int(cnt[0])
    xy = xy[0, :n].cpu().numpy().astype(np.int64)
    order = np.argsort(xy[:, 1] * map.shape[1] + xy[:, 0], kind="stable")
    return xy[order][:, ::-1]


def get_point_coordinates(map, scale_value=1., num_points=1000, threshold=-1, order_coord='xysr'):
    """test_utils.py:56-72"""
    idx = find_index_higher_scores(map, num_points=num_points, threshold=threshold)
    if len(idx) == 0:
        return np.asarray([])
    ys, xs = idx[:, 0], idx[:, 1]
    first, second = (xs, ys) if order_coord == 'xysr' else (ys, xs)
    out = np.empty((len(idx), 4), np.float64)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3] = first, second, scale_value, map[ys, xs]
    return out


# ------------------------------------------------------------------------------- greedy path
K_MAX = 16384          # capacity of one call of the device top-k (shared-memory sort of balf_greedy_nms_topk)


def _max_keep(h, w, r):
    return -(-h // (r + 1)) * -(-w // (r + 1))


def _keep_capacity(h, w, r):
    """the greedy NMS keeps at most one point per (r + 1) x (r + 1) cell; beyond K_MAX the result could be truncated"""
    return min(max(_max_keep(h, w, r), 1), K_MAX)


def _check_not_truncated(n, h, w, r):
    if n >= K_MAX and _max_keep(h, w, r) > K_MAX:
        raise ValueError("greedy NMS with radius %d on a %dx%d map can keep more than %d points; this drop-in returns at "
                         "most %d per call and refuses to truncate silently (the reference has no limit): use a larger "
                         "nms_size or split the map" % (r, h, w, K_MAX, K_MAX))


def get_points_direct_from_score_map(heatmap, conf_thresh=0.015, nms_size=15, subpixel=True, patch_size=5,
                                     scale_value=1., order_coord='xysr'):
    """test_utils.py:97-128 -- threshold, greedy NMS, optional sub-pixel; rows sorted by score desc."""
    h, w = heatmap.shape[0], heatmap.shape[1]
    k = _keep_capacity(h, w, nms_size)
    xy, sc, dxdy, cnt = _capi.greedy_nms_topk(_up(heatmap), k, border=0, thr=conf_thresh, radius=nms_size,
                                              subpixel_ps=patch_size if subpixel else 0)
    n = int(cnt[0])
    _check_not_truncated(n, h, w, nms_size)
    if n == 0:
        return np.zeros((0, 4))
    pts = xy[0, :n].cpu().numpy().astype(np.float64)
    if subpixel:
        pts = pts + dxdy[0, :n].cpu().numpy().astype(np.float64)
    first, second = (pts[:, 0], pts[:, 1]) if order_coord == 'xysr' else (pts[:, 1], pts[:, 0])
    return np.stack([first, second, np.full(n, float(scale_value)), sc[0, :n].cpu().numpy().astype(np.float64)], 1)


def nms_fast(in_corners, H, W, dist_thresh):
    """test_utils.py:130-168 -- in_corners [3,N] (x, y, score) -> (out [3,M] = the kept columns of ``in_corners`` itself
    (unrounded x, y and the original score values, :161), score-descending; indices into the input).  Candidates must round
    to distinct pixels (the pipeline's always do; the reference keeps the last-written index of a shared pixel)."""
    n = in_corners.shape[1]
    if n == 0:
        return np.zeros((3, 0)).astype(int), np.zeros(0).astype(int)
    rc = in_corners[:2].round().astype(int)
    if n == 1:
        return np.vstack((rc, in_corners[2])).reshape(3, 1), np.zeros((1)).astype(int)
    flat = rc[1] * W + rc[0]
    if len(np.unique(flat)) != n:
        raise NotImplementedError("nms_fast: several candidates round to the same pixel")
    dense = np.full((H, W), -np.inf, np.float32)
    dense.reshape(-1)[flat] = in_corners[2]
    k = _keep_capacity(H, W, dist_thresh)
    xy, _, _, cnt = _capi.greedy_nms_topk(_up(dense), k, border=0, thr=-3.0e38, radius=dist_thresh)
    m = int(cnt[0])
    _check_not_truncated(m, H, W, dist_thresh)
    kept = xy[0, :m].cpu().numpy().astype(np.int64)
    lut = np.full(H * W, -1, np.int64)
    lut[flat] = np.arange(n)
    inds = lut[kept[:, 1] * W + kept[:, 0]]
    return in_corners[:, inds], inds


def soft_argmax_points(pts, heatmap, patch_size=5):
    """test_utils.py:170-182 -- pts [3,N] (x, y, score) -> same with sub-pixel x, y."""
    out = np.array(pts, dtype=np.float64, copy=True)
    if out.shape[1] == 0:
        return out
    hm = np.asarray(heatmap.detach().cpu().numpy() if isinstance(heatmap, torch.Tensor) else heatmap).squeeze()
    xy = torch.from_numpy(out[:2].T.astype(np.int32)).contiguous().to(_dev())[None]
    dxdy = _capi.subpixel_refine(_up(hm), xy, patch_size)[0].cpu().numpy().astype(np.float64)
    out[:2] += dxdy.T
    return out
