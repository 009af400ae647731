"""B200-native drop-in for the hot functions of the reference's ``balf/benchmark_test/repeatability_tools.py``
(SURVEY.md section 8 f3): ``compute_repeatability`` (:379-490), ``compute_resize_repeatability`` (:516-614) and
``box_nms`` (:227-255).  Same names, arguments and result dictionaries; NumPy in / NumPy out, the O(N1 N2) work runs in
``libbalf_b200.so`` (``csrc/metrics.cu``, ``csrc/nms.cu``).  Ties in the reference's (unstable) sorts resolve in index
order here.  There is no CPU path."""
import numpy as np
import torch

from .. import _capi
from .geometry_tools import _dev


def compute_repeatability(src_indexes, dst_indexes, overlap_err=0.4, eps=1e-6, dist_match_thresh=3, radious_size=30.,
                          device=None):
    src = np.ascontiguousarray(np.asarray(src_indexes, np.float64).reshape(-1, 4))
    dst = np.ascontiguousarray(np.asarray(dst_indexes, np.float64).reshape(-1, 4))
    n1, n2 = len(src), len(dst)
    if n1 == 0 or n2 == 0:                       # the reference's loops do not run; it divides by zero points (nan + warning)
        with np.errstate(divide="ignore", invalid="ignore"):
            nan = np.asarray(0.0) / np.asarray(0, float) * 100.0
        return {'rep_single_scale': nan, 'rep_multi_scale': nan, 'num_points_single_scale': 0, 'num_points_multi_scale': 0,
                'error_overlap_single_scale': 0.0, 'error_overlap_multi_scale': 0.0, 'total_num_points': 0,
                'correspondences': np.asarray([]), 'possible_matches': 0, 'correspondences_m': np.asarray([])}
    dev = _dev(device)
    sc, cs, cm, ov = _capi.compute_repeatability(torch.from_numpy(src).to(dev), torch.from_numpy(dst).to(dev), overlap_err, eps,
                                                 dist_match_thresh, radious_size)
    sc = sc.cpu().numpy()
    if int(ov.item()):
        raise RuntimeError("compute_repeatability: more than 64 candidate pairs per point reached the overlap threshold "
                           "(degenerate input: many coincident points); the workspace limit is documented in include/balf_b200.h")
    fs, fm = int(sc[2]), int(sc[3])
    return {'rep_single_scale': sc[0], 'rep_multi_scale': sc[1], 'num_points_single_scale': fs,
            'num_points_multi_scale': fm, 'error_overlap_single_scale': float(sc[4]),
            'error_overlap_multi_scale': float(sc[5]), 'total_num_points': int(sc[6]),
            'correspondences': cs[:fs].cpu().numpy().astype(np.int64) if fs else np.asarray([]),
            'possible_matches': int(sc[7]),
            'correspondences_m': cm[:fm].cpu().numpy().astype(np.int64) if fm else np.asarray([])}


def compute_resize_repeatability(keypoints, warped_keypoints, h, shape_src, shape_dst, keep_k_points=1000, distance_thresh=5,
                                 device=None):
    """The reference overwrites ``keypoints[:, :2]`` in place (:560-561); this implementation leaves its inputs alone."""
    dev = _dev(device)
    kp = torch.from_numpy(np.ascontiguousarray(np.asarray(keypoints, np.float64).reshape(-1, 3))).to(dev)
    wkp = torch.from_numpy(np.ascontiguousarray(np.asarray(warped_keypoints, np.float64).reshape(-1, 3))).to(dev)
    o = _capi.resize_repeatability(kp, wkp, h, shape_src, shape_dst, keep_k_points, distance_thresh).cpu().numpy()
    return {'repeatability': float(o[0]), 'localization_err': float(o[1]) if o[1] >= 0 else -1,
            'common_src_num': int(o[2]), 'common_dst_num': int(o[3]), 'rep_src_num': int(o[4]), 'rep_dst_num': int(o[5])}


def box_nms(prob, size=4, iou=0.1, min_prob=0.015, keep_top_k=-1):
    """repeatability_tools.py:227-255.  prob: torch tensor [1,H,W] (any device; the reference moves the boxes to the GPU
    itself) -> tensor [1,H,W] on prob's device."""
    assert prob.shape[0] == 1 and len(prob.shape) == 3
    dev = prob.device if prob.is_cuda else _dev(None)
    return _capi.box_nms_map(prob.to(dev), size, iou, min_prob, keep_top_k).to(prob.device)
