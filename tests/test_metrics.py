"""SURVEY.md section 8 f3 -- repeatability metrics and homography helpers.
CPU: oracle/metrics.py against the reference-generated vectors (tests/golden/r2_metrics.npz) and the OpenCV restatement
against cv2 itself.  GPU: the CUDA kernels (csrc/metrics.cu, through the C-ABI and the drop-in modules) against both."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import metrics as om
from oracle.make_golden import metric_inputs

CASES = [("unit", 0.4), ("unit", 0.2), ("ms", 0.4), ("ms", 0.2), ("raw", 0.4), ("raw", 0.2)]
SCALARS = ["rep_single_scale", "rep_multi_scale", "num_points_single_scale", "num_points_multi_scale",
           "error_overlap_single_scale", "error_overlap_multi_scale", "total_num_points", "possible_matches"]


def _case(g, tag):
    Hm, shape, src, dst, src_ms, dst_ms = metric_inputs()
    return {"unit": (src, g["dst_to_src"]), "ms": (src_ms, g["dst_to_src_ms"]), "raw": (src, dst)}[tag]


def _check_rep(res, g, tag, oe, err_tol):
    pre = "rep_%s_oe%d_" % (tag, int(oe * 10))
    want = g[pre + "scalars"]
    got = np.array([res[k] for k in SCALARS], np.float64)
    np.testing.assert_array_equal(got[[2, 3, 6, 7]], want[[2, 3, 6, 7]])            # counts: exact
    np.testing.assert_allclose(got[[0, 1]], want[[0, 1]], rtol=1e-14)
    np.testing.assert_allclose(got[[4, 5]], want[[4, 5]], rtol=0, atol=err_tol)
    np.testing.assert_array_equal(np.asarray(res["correspondences"]).reshape(-1, 2), g[pre + "corr"])
    np.testing.assert_array_equal(np.asarray(res["correspondences_m"]).reshape(-1, 2), g[pre + "corr_m"])


# ----------------------------------------------------------------------------- CPU: the oracle is pinned
@pytest.mark.parametrize("tag,oe", CASES)
def test_oracle_compute_repeatability_vs_reference(tag, oe):
    g = load_golden("r2_metrics.npz")
    a, b = _case(g, tag)
    _check_rep(om.compute_repeatability(a, b, overlap_err=oe), g, tag, oe, 1e-15)


def test_oracle_homography_points_vs_reference():
    g = load_golden("r2_metrics.npz")
    Hm, _, src, dst, src_ms, dst_ms = metric_inputs()
    np.testing.assert_allclose(om.apply_homography_to_points(dst, Hm), g["dst_to_src"], rtol=1e-12)
    np.testing.assert_allclose(om.apply_homography_to_points(dst_ms, Hm), g["dst_to_src_ms"], rtol=1e-12)
    assert om.apply_homography_to_points(np.zeros((0, 4)), Hm).size == 0


def test_oracle_resize_repeatability_vs_reference():
    g = load_golden("r2_metrics.npz")
    Hm, (hs, ws), src, dst, _, _ = metric_inputs()
    kp = np.stack([src[:, 1], src[:, 0], src[:, 3]], 1)
    wkp = np.stack([dst[:, 1], dst[:, 0], dst[:, 3]], 1)
    for k, thr in ((1000, 5), (150, 3), (50, 1)):
        r = om.compute_resize_repeatability(kp, wkp, np.linalg.inv(Hm), (hs, ws), (hs, ws), k, thr)
        got = np.array([r["repeatability"], r["localization_err"], r["common_src_num"], r["common_dst_num"],
                        r["rep_src_num"], r["rep_dst_num"]], np.float64)
        np.testing.assert_allclose(got, g["resize_k%d_t%d" % (k, thr)], rtol=1e-13)


def test_oracle_common_region_masks_vs_reference_and_cv2():
    g = load_golden("r2_metrics.npz")
    _, (hs, ws), *_ = metric_inputs()
    for i in range(2):
        ms, md = om.create_common_region_masks(g["mask_H%d" % i], (hs, ws, 3), (hs + 16, ws - 24, 3))
        np.testing.assert_array_equal(ms.astype(np.uint8), g["mask_src_%d" % i])
        np.testing.assert_array_equal(md.astype(np.uint8), g["mask_dst_%d" % i])
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for _ in range(4):                       # the restated warpPerspective itself, on non-binary images
        Hx = np.eye(3) + rng.normal(0, 0.03, (3, 3)) * np.array([[1, 1, 200], [1, 1, 200], [1e-3, 1e-3, 0]])
        img = rng.random((97, 131))
        want = cv2.warpPerspective(img, Hx, (150, 83))
        np.testing.assert_allclose(om.warp_perspective_linear(img, Hx, (150, 83)), want, rtol=0, atol=1e-12)


# ----------------------------------------------------------------------------- GPU: the kernels
@pytest.mark.gpu
@pytest.mark.parametrize("tag,oe", CASES)
def test_gpu_compute_repeatability(tag, oe):
    from balf_b200.benchmark_test import repeatability_tools as rt
    g = load_golden("r2_metrics.npz")
    a, b = _case(g, tag)
    # overlaps come from CUDA's acos / sin instead of libm's: 1 - overlap sums agree to a few ulp
    _check_rep(rt.compute_repeatability(a, b, overlap_err=oe), g, tag, oe, 1e-12)


@pytest.mark.gpu
def test_gpu_compute_repeatability_larger_random_vs_oracle():
    from balf_b200.benchmark_test import repeatability_tools as rt
    rng = np.random.default_rng(11)
    a = np.stack([rng.uniform(0, 600, 1500), rng.uniform(0, 400, 1500), rng.uniform(0.5, 4, 1500), rng.random(1500)], 1)
    b = np.concatenate([a[:900, :2] + rng.normal(0, 2.0, (900, 2)), rng.uniform(0, 400, (400, 2))], 0)
    b = np.concatenate([b, rng.uniform(0.5, 4, (1300, 1)), rng.random((1300, 1))], 1)
    want = om.compute_repeatability(a, b)
    got = rt.compute_repeatability(a, b)
    for k in SCALARS:
        np.testing.assert_allclose(got[k], want[k], rtol=1e-12, atol=1e-12)
    np.testing.assert_array_equal(got["correspondences"], want["correspondences"])
    np.testing.assert_array_equal(got["correspondences_m"], want["correspondences_m"])
    empty = rt.compute_repeatability(np.zeros((0, 4)), b)
    assert empty["num_points_single_scale"] == 0 and empty["possible_matches"] == 0


@pytest.mark.gpu
def test_gpu_homography_points_and_masks():
    from balf_b200.benchmark_test import geometry_tools as gt
    g = load_golden("r2_metrics.npz")
    Hm, (hs, ws), src, dst, src_ms, dst_ms = metric_inputs()
    np.testing.assert_allclose(gt.apply_homography_to_points(dst, Hm), g["dst_to_src"], rtol=1e-12)
    np.testing.assert_allclose(gt.apply_homography_to_points(dst_ms, Hm), g["dst_to_src_ms"], rtol=1e-12)
    assert gt.apply_homography_to_points(np.zeros((0, 4)), Hm).size == 0
    for i in range(2):
        ms, md = gt.create_common_region_masks(g["mask_H%d" % i], (hs, ws, 3), (hs + 16, ws - 24, 3))
        np.testing.assert_array_equal(ms.astype(np.uint8), g["mask_src_%d" % i])
        np.testing.assert_array_equal(md.astype(np.uint8), g["mask_dst_%d" % i])
    rng = np.random.default_rng(3)
    for _ in range(3):                       # larger frames against the oracle restatement
        Hx = np.eye(3) + rng.normal(0, 0.02, (3, 3)) * np.array([[1, 1, 300], [1, 1, 300], [2e-4, 2e-4, 0]])
        ms, md = gt.create_common_region_masks(Hx, (480, 640, 3), (470, 700, 3))
        ws_, wd_ = om.create_common_region_masks(Hx, (480, 640, 3), (470, 700, 3))
        np.testing.assert_array_equal(ms, ws_)
        np.testing.assert_array_equal(md, wd_)


@pytest.mark.gpu
def test_gpu_resize_repeatability():
    from balf_b200.benchmark_test import repeatability_tools as rt
    g = load_golden("r2_metrics.npz")
    Hm, (hs, ws), src, dst, _, _ = metric_inputs()
    kp = np.stack([src[:, 1], src[:, 0], src[:, 3]], 1)
    wkp = np.stack([dst[:, 1], dst[:, 0], dst[:, 3]], 1)
    kp0 = kp.copy()
    for k, thr in ((1000, 5), (150, 3), (50, 1)):
        r = rt.compute_resize_repeatability(kp, wkp, np.linalg.inv(Hm), (hs, ws), (hs, ws), k, thr)
        got = np.array([r["repeatability"], r["localization_err"], r["common_src_num"], r["common_dst_num"],
                        r["rep_src_num"], r["rep_dst_num"]], np.float64)
        np.testing.assert_allclose(got, g["resize_k%d_t%d" % (k, thr)], rtol=1e-12)
    np.testing.assert_array_equal(kp, kp0)
    r = rt.compute_resize_repeatability(np.zeros((0, 3)), wkp, np.linalg.inv(Hm), (hs, ws), (hs, ws))
    want = om.compute_resize_repeatability(np.zeros((0, 3)), wkp, np.linalg.inv(Hm), (hs, ws), (hs, ws))
    assert r["repeatability"] == want["repeatability"] and r["localization_err"] == want["localization_err"]
