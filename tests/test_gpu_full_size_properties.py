"""BASELINE.json configs[1] at full size (batch 64, 480x640, top-2048) through size-independent properties of the domain
(the CPU oracle needs ~4 s per image at this size, so the batch is checked by what must hold for ANY correct result):
determinism, per-image independence, ordering, bounds, the defining property of each NMS (windowed: every keypoint is the
maximum of its 15x15 window of the border-masked map; greedy: no two keypoints within Chebyshev distance 15, and every
above-threshold pixel is within 15 of a kept keypoint that is at least as strong), and one image against the oracle."""
import copy

import numpy as np
import pytest
import torch

from conftest import synth_u8

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
B, H, W, K = 64, 480, 640, 2048


@pytest.fixture(scope="module")
def batch():
    return torch.from_numpy(np.stack([synth_u8(H, W, 1000 + i)[:, :, :1] for i in range(B)]))


def _score_maps(det, u8):
    import balf_b200._capi as capi
    x, (top, left) = capi.preprocess_u8(u8)
    with torch.inference_mode():
        prob = det(x)["prob"]
    return prob[:, top:top + H, left:left + W].contiguous()


def test_windowed_full_batch_properties(detector, batch):
    from balf_b200.configs import config
    from balf_b200.demo import demo_match
    det = copy.deepcopy(detector).to(DEV).eval()
    args = config.default_test_args(sub_pixel=False, num_features=K)
    u8 = batch.to(DEV)
    xy, sc, _, cnt = demo_match.detect_batch_device(args, u8, det, "windowed")
    xy2, sc2, _, cnt2 = demo_match.detect_batch_device(args, u8, det, "windowed")
    assert torch.equal(xy, xy2) and torch.equal(sc, sc2) and torch.equal(cnt, cnt2)           # run-to-run identical
    for i in (0, 17, 63):                                                                       # image i alone == image i in the batch
        xi, si, _, ci = demo_match.detect_batch_device(args, u8[i:i + 1], det, "windowed")
        assert int(ci[0]) == int(cnt[i]) and torch.equal(xi[0], xy[i]) and torch.equal(si[0], sc[i])
    prob = _score_maps(det, u8)
    masked = torch.zeros_like(prob)
    b = args.border_size
    masked[:, b:H - b, b:W - b] = prob[:, b:H - b, b:W - b]
    wmax = torch.nn.functional.max_pool2d(masked[:, None], 15, stride=1, padding=7)[:, 0]     # checker only (torch op on the result)
    for i in range(B):
        n = int(cnt[i])
        assert 0 < n <= K
        p = xy[i, :n].long()
        s = sc[i, :n]
        assert (p[:, 0] >= b).all() and (p[:, 0] < W - b).all() and (p[:, 1] >= b).all() and (p[:, 1] < H - b).all()
        assert (s[:-1] >= s[1:]).all()                                                          # score-descending
        assert torch.equal(s, prob[i, p[:, 1], p[:, 0]])                                        # scores are the map's values
        assert torch.equal(s, wmax[i, p[:, 1], p[:, 0]])                                        # each is its window's maximum
        lin = p[:, 1] * W + p[:, 0]
        assert lin.unique().numel() == n
        ties = s[:-1] == s[1:]
        assert (lin[:-1][ties] < lin[1:][ties]).all()                                           # canonical tie rule


def test_greedy_full_batch_properties(detector, batch):
    from balf_b200.configs import config
    from balf_b200.demo import demo_match
    det = copy.deepcopy(detector).to(DEV).eval()
    args = config.default_test_args(sub_pixel=False, num_features=K)
    u8 = batch[:16].to(DEV)
    xy, sc, _, cnt = demo_match.detect_batch_device(args, u8, det, "greedy")
    prob = _score_maps(det, u8)
    b, r, thr = args.border_size, args.nms_size, np.float32(args.heatmap_confidence_threshold)
    for i in range(16):
        n = int(cnt[i])
        assert 0 < n < K                                       # top-k does not bind here (about 850 keypoints per image)
        p = xy[i, :n].long()
        s = sc[i, :n]
        assert (s[:-1] >= s[1:]).all() and (s >= float(thr)).all()
        assert torch.equal(s, prob[i, p[:, 1], p[:, 0]])
        d = (p[:, None, :] - p[None, :, :]).abs().amax(-1)
        d.fill_diagonal_(10 ** 6)
        assert int(d.min()) > r                                # no two kept keypoints inside each other's exclusion box
        # maximality: every candidate pixel is covered by a kept keypoint that is at least as strong
        kept = torch.zeros(H, W, device=DEV)
        kept[p[:, 1], p[:, 0]] = s
        cover = torch.nn.functional.max_pool2d(kept[None, None], 2 * r + 1, stride=1, padding=r)[0, 0]
        cand = torch.zeros(H, W, dtype=torch.bool, device=DEV)
        cand[b:H - b, b:W - b] = prob[i, b:H - b, b:W - b] >= float(thr)
        assert bool((cover[cand] >= prob[i][cand]).all())


def test_one_full_size_image_against_the_oracle(detector, detector_sd, batch):
    from balf_b200.configs import config
    from balf_b200.demo import demo_match
    from oracle import pipeline
    det = copy.deepcopy(detector).to(DEV).eval()
    det.precision = "fp32"
    args = config.default_test_args(sub_pixel=False, num_features=K)
    im = np.repeat(batch[5].numpy(), 3, axis=2)
    xy, sc, _, cnt = demo_match.detect_batch_device(args, batch[5:6].to(DEV), det, "windowed")
    want = pipeline.detect_windowed(detector_sd, im, 15, 15, K)
    n = int(cnt[0])
    assert n == len(want) == K
    # each side on its own score map (fp32 FFMA path vs the CPU oracle: rel 2e-6): near-equal scores may swap places and
    # the k-th value cut may pick a different last few, so the comparison is on the set (bit-exactness is tested on
    # identical score maps in test_gpu_postproc.py)
    got = set(map(tuple, xy[0, :n].cpu().numpy().tolist()))
    ref = set(map(tuple, want[:, :2].astype(np.int32).tolist()))
    assert len(got & ref) >= 0.99 * K, len(got & ref)
    np.testing.assert_allclose(np.sort(sc[0, :n].cpu().numpy())[::-1][:K // 2], np.sort(want[:, 3])[::-1][:K // 2], rtol=2e-5)
