"""Parity of the tensor-core detector path (precision 'tf32': balf_b200/csrc/detector_tc.cuh, tcgen05
kind::tf32 with fp32 accumulation in TMEM) with the reference's golden vectors and the CPU oracle.

Bound (BASELINE.json north_star): score maps within rel <= 1e-3 for fp32-accumulate paths, and
>= 99 % keypoint agreement end to end."""
import copy

import numpy as np
import pytest
import torch

from conftest import load_golden, synth_u8
from oracle import detector as odet
from oracle import pipeline, postproc_c

pytestmark = pytest.mark.gpu
TF32_RTOL = 1e-3


@pytest.fixture(scope="module")
def det_tc(detector):
    d = copy.deepcopy(detector).to("cuda:0").eval()
    d.precision = "tf32"
    return d


def run(det, x):
    with torch.inference_mode():
        o = det(x.to("cuda:0"))
    torch.cuda.synchronize()
    return o["logits"].cpu(), o["prob"].cpu()


def test_golden_small(det_tc):
    g = load_golden("detector.npz")
    x = torch.rand(1, 3, 128, 192, generator=torch.Generator().manual_seed(1234))
    logits, prob = run(det_tc, x)
    np.testing.assert_allclose(prob[0].numpy(), g["prob_128x192"], rtol=TF32_RTOL)
    np.testing.assert_allclose(logits[0].numpy(), g["logits_128x192"], atol=1.5e-3)
    x2 = torch.rand(2, 3, 64, 128, generator=torch.Generator().manual_seed(77))
    logits, prob = run(det_tc, x2)
    np.testing.assert_allclose(prob.numpy(), g["prob_b2_64x128"], rtol=TF32_RTOL)


def test_golden_512x640(det_tc):
    g = load_golden("detector.npz")
    x = torch.rand(1, 3, 512, 640, generator=torch.Generator().manual_seed(1234))
    logits, prob = run(det_tc, x)
    p = prob[0].numpy()
    np.testing.assert_allclose(p[::8, ::8], g["prob_512x640_sub8"], rtol=TF32_RTOL)
    np.testing.assert_allclose(p[255], g["prob_512x640_row255"], rtol=TF32_RTOL)
    assert abs(p.astype(np.float64).sum() - 5044.952016152) < 0.05


def test_batches_units_and_determinism(det_tc, detector_sd):
    # odd unit counts (64x64 -> a single 64-token unit at the last stage), batches that cross the
    # internal chunk of 16 images, per-image independence and run-to-run bit reproducibility
    import balf_b200._capi as capi
    xb = torch.rand(19, 3, 64, 64, generator=torch.Generator().manual_seed(3))
    _, pb = run(det_tc, xb)                      # automatic chunk: one pass of 19 images
    _, pb2 = run(det_tc, xb)
    np.testing.assert_array_equal(pb.numpy(), pb2.numpy())
    capi.debug_set(1, 16)                        # pin the internal pass to 16 images: 19 = 16 + 3 crosses it
    try:
        _, pb3 = run(det_tc, xb)
        np.testing.assert_array_equal(pb.numpy(), pb3.numpy())
        for i in (0, 7, 15, 16, 18):
            _, pi = run(det_tc, xb[i:i + 1])
            np.testing.assert_array_equal(pb[i].numpy(), pi[0].numpy())
    finally:
        capi.debug_set(1, 0)
    with torch.inference_mode():
        o = odet.detector_forward(detector_sd, xb)
    np.testing.assert_allclose(pb.numpy(), o["prob"].numpy(), rtol=TF32_RTOL)
    x3 = torch.rand(3, 3, 192, 64, generator=torch.Generator().manual_seed(4))
    _, p3 = run(det_tc, x3)
    with torch.inference_mode():
        o3 = odet.detector_forward(detector_sd, x3)
    np.testing.assert_allclose(p3.numpy(), o3["prob"].numpy(), rtol=TF32_RTOL)


def test_keypoint_agreement(det_tc, detector_sd):
    """each side on its own score map: >= 99 % of the oracle's keypoints (integer coordinates)."""
    from balf_b200.configs import config
    from balf_b200.demo import demo_match
    g = load_golden("detect.npz")
    args = config.default_test_args(sub_pixel=False)
    # the small golden images hold ~50 keypoints each (one flipped local maximum = 2 %), so the 99 % bound is taken
    # over their union and each image may lose at most one keypoint
    n_inter = n_ref = 0
    for h, w, seed in ((121, 187, 5), (128, 192, 6)):
        im = synth_u8(h, w, seed)
        got = demo_match.detect(args, im, det_tc, "cuda:0")
        ref = g["detect_%dx%d" % (h, w)]                     # the reference's own detect() output
        inter = set(map(tuple, got[:, :2])) & set(map(tuple, ref[:, :2]))
        assert len(inter) >= len(ref) - 1, (len(inter), len(ref))
        n_inter += len(inter)
        n_ref += len(ref)
    assert n_inter >= 0.99 * n_ref, (n_inter, n_ref)
    # 480x640.  The demo path (demo_match.detect: greedy nms_fast) runs the detector in its DEFAULT precision ('auto' ->
    # 'f16x3' for this path: the tensor-core kernels on fp16 hi + lo operand pairs, fp32-class score maps, see
    # MLP_MA_DECODER.resolve_precision): >= 99 % on EVERY image, all six seeds of scripts/tc_precision.py (measured: 100 %
    # against the fp32 path on all six).  Opting into tf32 on the greedy path is allowed but documented as below the target:
    # random-init score maps are nearly flat (0.010-0.023) and the greedy suppression chains amplify 1e-4 score
    # perturbations (measured 0.975-0.997 per image, 0.989 over six seeds); the windowed extraction -- the throughput /
    # benchmark path, whose default IS tf32 -- keeps >= 99 % on every image.
    d_auto = copy.deepcopy(det_tc)
    d_auto.precision = "auto"
    n_inter = n_ref = 0
    for seed in (1234, 1, 2, 3, 4, 5):
        im = synth_u8(480, 640, seed)
        want = pipeline.detect(args, detector_sd, im, nms=postproc_c.greedy_nms)
        ws = set(map(tuple, want[:, :2]))
        got = demo_match.detect(args, im, d_auto, "cuda:0")
        inter = set(map(tuple, got[:, :2])) & ws
        assert len(inter) >= 0.99 * len(want), ("default precision, greedy", seed, len(inter), len(want))
        if seed in (1234, 1, 2):
            got = demo_match.detect(args, im, det_tc, "cuda:0")                 # explicit tf32 opt-in
            inter = set(map(tuple, got[:, :2])) & ws
            assert len(inter) >= 0.97 * len(want), (seed, len(inter), len(want))
            n_inter += len(inter)
            n_ref += len(want)
        # windowed path (validation extraction), top-2048, default precision of that path = tf32
        xy, sc, _, cnt = demo_match.detect_batch_device(config.default_test_args(sub_pixel=False), torch.from_numpy(im[None, :, :, :1].copy()).to("cuda:0"),
                                                        d_auto, nms="windowed")
        wantw = pipeline.detect_windowed(detector_sd, im, 15, 15, 2048)
        gotw = set(map(tuple, xy[0, :int(cnt[0])].cpu().numpy().tolist()))
        ww = set(map(tuple, wantw[:, :2].astype(int).tolist()))
        assert len(gotw & ww) >= 0.99 * len(ww), (seed, len(gotw & ww), len(ww))
    assert n_inter >= 0.985 * n_ref, (n_inter, n_ref)
    assert d_auto.resolve_precision("windowed") == "tf32" and d_auto.resolve_precision("greedy") == "f16x3" \
        and d_auto.resolve_precision() == "f16x3" and det_tc.resolve_precision("greedy") == "tf32"


def test_large_shapes_vs_reference_maps(det_tc):
    """BASELINE.json configs[2] / [3] shapes against the REFERENCE's own score maps (tests/golden/r2_detector_large.npz:
    900x1200 -> 960x1216 padded, 1024x1024, written by oracle/make_golden.py from /root/reference): tolerance of the
    north star for the tensor-core path, 2e-5 for the fp32 path and for the split-precision tensor path ('f16x3');
    per-image independence across the internal pass boundary and bit reproducibility (both tensor-core precisions)."""
    import balf_b200._capi as capi
    g = load_golden("r2_detector_large.npz")
    d32 = copy.deepcopy(det_tc)
    d32.precision = "fp32"
    dx3 = copy.deepcopy(det_tc)
    dx3.precision = "f16x3"
    for h, w, seeds in ((900, 1200, (1234, 1235)), (1024, 1024, (1234,))):
        ims = np.stack([synth_u8(h, w, s)[:, :, :1] for s in seeds])
        x, _ = capi.preprocess_u8(torch.from_numpy(ims).to("cuda:0"))
        assert tuple(x.shape[2:]) == tuple(g["pad_shape_%dx%d_s%d" % (h, w, seeds[0])][:2])
        _, p_tc = run(det_tc, x)
        _, p_32 = run(d32, x)
        _, p_x3 = run(dx3, x)
        for i, s in enumerate(seeds):
            key = "%dx%d_s%d" % (h, w, s)
            for p, tol in ((p_tc, TF32_RTOL), (p_32, 2e-5), (p_x3, 2e-5)):
                pi = p[i].numpy()
                np.testing.assert_allclose(pi[::8, ::8], g["prob_sub8_" + key], rtol=tol)
                np.testing.assert_allclose(pi[pi.shape[0] // 2 - 1], g["prob_row_" + key], rtol=tol)
                assert abs(pi.astype(np.float64).sum() - g["prob_stats_" + key][0]) < (0.5 if tol > 1e-4 else 0.02)
        capi.debug_set(1, 1)                         # one image per internal pass
        try:
            _, p_one = run(det_tc, x)
            _, p_one3 = run(dx3, x)
        finally:
            capi.debug_set(1, 0)
        np.testing.assert_array_equal(p_tc.numpy(), p_one.numpy())
        np.testing.assert_array_equal(p_x3.numpy(), p_one3.numpy())
