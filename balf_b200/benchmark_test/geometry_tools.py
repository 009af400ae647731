"""B200-native drop-in for the hot functions of the reference's ``balf/benchmark_test/geometry_tools.py``
(SURVEY.md section 8 f3): same names, NumPy in / NumPy out, the arithmetic runs in ``libbalf_b200.so``
(``csrc/metrics.cu``) on the current CUDA device.  There is no CPU path."""
import numpy as np
import torch

from .. import _capi


def _dev(device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("balf_b200 has no CPU path: a CUDA device is required")
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def create_common_region_masks(h_dst_2_src, shape_src, shape_dst, device=None):
    """geometry_tools.py:7-27 -> (mask_src, mask_dst) float64 arrays of 0.0 / 1.0."""
    ms, md = _capi.common_region_masks(h_dst_2_src, shape_src, shape_dst, _dev(device), border=15)
    return ms.cpu().numpy().astype(np.float64), md.cpu().numpy().astype(np.float64)


def apply_homography_to_points(points, h, device=None):
    """geometry_tools.py:43-64: [n,4] (x, y, radius, score) -> [n,4]; an empty input returns ``np.asarray([])`` as the
    reference does."""
    pts = np.asarray(points, np.float64)
    if pts.size == 0:
        return np.asarray([])
    out = _capi.apply_homography_to_points(torch.from_numpy(np.ascontiguousarray(pts.reshape(-1, 4))).to(_dev(device)), h)
    return out.cpu().numpy()
