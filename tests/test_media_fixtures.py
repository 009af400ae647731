"""BASELINE.json configs[0]: the reference's only shipped inputs, media/im1.jpg and media/im2.jpg (768x480), through
the reference path.  tests/golden/r2_media.npz holds the images as decoded by PIL (RGB and convert('L'), demo_match.load_im
:13-19) and the reference's own detect() output on them (oracle/make_golden.py r2, run in the build container).
CPU: the oracle against those vectors; GPU: the CUDA path against them."""
import copy

import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import multiscale as oms
from oracle import pipeline, postproc_c, weights


def test_oracle_luma_is_pil_convert_L():
    g = load_golden("r2_media.npz")
    for name in ("im1", "im2"):
        np.testing.assert_array_equal(oms.rgb_to_gray(g["rgb_" + name]), g["gray_" + name])
        assert g["rgb_" + name].shape == (480, 768, 3)


def test_oracle_detect_on_media_vs_reference():
    """SURVEY.md 8c anchor: 332 100 candidates at thr 0.001, nms_fast keeps 1049 on im1 (1005 on im2)."""
    g = load_golden("r2_media.npz")
    sd = weights.detector_state_dict(0)
    args = pipeline.default_args(sub_pixel=False)
    for name, n in (("im1", 1049), ("im2", 1005)):
        got = pipeline.detect(args, sd, g["rgb_" + name], nms=postproc_c.greedy_nms)
        ref = g["detect_" + name]
        assert len(ref) == n and int(g["candidates_" + name][0]) == 332100
        inter = set(map(tuple, got[:, :2])) & set(map(tuple, ref[:, :2]))
        assert len(got) == n and len(inter) >= n - 2, (name, len(got), len(inter))


@pytest.mark.gpu
def test_gpu_luma_and_detect_on_media(detector):
    import balf_b200._capi as capi
    from balf_b200.configs import config
    from balf_b200.demo import demo_match
    g = load_golden("r2_media.npz")
    dev = torch.device("cuda:0")
    det = copy.deepcopy(detector).to(dev).eval()
    args = config.default_test_args(sub_pixel=False)
    for name, n in (("im1", 1049), ("im2", 1005)):
        rgb = g["rgb_" + name]
        gray = capi.rgb_to_gray(torch.from_numpy(rgb).to(dev)).cpu().numpy()
        np.testing.assert_array_equal(gray, g["gray_" + name])                      # PIL convert('L'), bit for bit
        ref = g["detect_" + name]
        got = demo_match.detect(args, rgb, det, dev)                                 # default precision of the demo path
        assert got.shape == (n, 3) and (got[:, 2] == 1.0).all()
        inter = set(map(tuple, got[:, :2])) & set(map(tuple, ref[:, :2]))
        assert len(inter) >= 0.99 * n, (name, len(inter))
        # the score map itself against the reference's (sub-sampled) map, every precision
        x, (top, left) = capi.preprocess_u8(torch.from_numpy(rgb[None]).to(dev))
        for prec, tol in (("fp32", 2e-5), ("f16x3", 2e-5), ("tf32", 1e-3)):
            with torch.inference_mode():
                prob = det(x, precision=prec)["prob"][0].cpu().numpy()
            np.testing.assert_allclose(prob[::8, ::8], g["prob_sub8_" + name], rtol=tol)
        # windowed extraction on the throughput path (tf32 default) keeps >= 99 % of the oracle's keypoints
        xy, sc, _, cnt = demo_match.detect_batch_device(args, torch.from_numpy(rgb[None]).to(dev), det, nms="windowed")
        want = pipeline.detect_windowed(weights.detector_state_dict(0), rgb, 15, 15, 2048)
        gw = set(map(tuple, xy[0, :int(cnt[0])].cpu().numpy().tolist()))
        ww = set(map(tuple, want[:, :2].astype(int).tolist()))
        assert len(gw & ww) >= 0.99 * len(ww)
