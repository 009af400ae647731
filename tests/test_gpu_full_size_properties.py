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


def _score_maps(det, u8, nms="windowed"):
    import balf_b200._capi as capi
    x, (top, left) = capi.preprocess_u8(u8)
    with torch.inference_mode():
        prob = det(x, precision=det.resolve_precision(nms))["prob"]      # the arithmetic detect_batch_device(nms) runs in
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
    prob = _score_maps(det, u8, "greedy")
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


def test_config3_pair_properties(detector):
    """BASELINE.json configs[2]: an HPatches-shaped pair (900x1200 -> 960x1216 padded), detector + greedy NMS (top-2048
    binds: ~2900 survivors) + level-1 patches + HardNet + SMNN, through properties: shapes, unit-norm descriptors, matches
    are mutual nearest neighbours passing the 0.99 ratio test under the kernel's own distance matrix, ordered by the first
    index, each index used once; the second image is the first plus +-2 grey levels of noise, so most keypoints match."""
    import balf_b200._capi as capi
    from balf_b200.configs import config
    from balf_b200.demo import demo_match
    from balf_b200.third_party.hardnet.hardnet_pytorch import HardNet
    det = copy.deepcopy(detector).to(DEV).eval()
    torch.manual_seed(0)
    hn = HardNet().eval().to(DEV)
    args = config.default_test_args()
    rgb1 = synth_u8(900, 1200, 1234)
    noise = np.random.default_rng(5).integers(-2, 3, rgb1.shape[:2])[..., None]
    rgb2 = np.clip(rgb1.astype(np.int64) + noise, 0, 255).astype(np.uint8)
    g1, g2 = rgb1[..., 0].copy(), rgb2[..., 0].copy()
    k1, d1 = demo_match.extract_features(args, rgb1, g1, det, hn, DEV)
    k2, d2 = demo_match.extract_features(args, rgb2, g2, det, hn, DEV)
    assert k1.shape == (2048, 2) and d1.shape == (2048, 128) and k2.shape == (2048, 2)
    np.testing.assert_allclose(np.linalg.norm(d1, axis=1), 1.0, atol=1e-4)
    assert (k1[:, 0] >= 15 - 2).all() and (k1[:, 0] <= 1200 - 15 + 2).all() and (k1[:, 1] >= 15 - 2).all() and (k1[:, 1] <= 900 - 15 + 2).all()
    t1, t2 = torch.from_numpy(d1).to(DEV), torch.from_numpy(d2).to(DEV)
    dist, ids, dm = capi.match_smnn(t1, t2, 0.99, want_dm=True)
    ids, dm = ids.cpu().numpy(), dm.cpu().numpy()
    assert len(ids) > 200
    assert (np.diff(ids[:, 0]) > 0).all() and len(np.unique(ids[:, 1])) == len(ids)
    rows, cols = dm[ids[:, 0]], dm[:, ids[:, 1]].T
    np.testing.assert_array_equal(rows.argmin(1), ids[:, 1])                     # nearest neighbour both ways
    np.testing.assert_array_equal(cols.argmin(1), ids[:, 0])
    r12 = np.sort(rows, 1)[:, 0] / np.sort(rows, 1)[:, 1]
    r21 = np.sort(cols, 1)[:, 0] / np.sort(cols, 1)[:, 1]
    assert (r12 <= 0.99).all() and (r21 <= 0.99).all()
    np.testing.assert_allclose(dist.cpu().numpy()[:, 0], np.maximum(r12, r21), rtol=1e-6)
    p1, p2 = demo_match.extract_matches(args, rgb1, g1, rgb2, g2, det, hn, DEV)
    assert p1.shape == p2.shape == (len(ids), 2)
    assert np.median(np.abs(p1 - p2).max(1)) < 1.0                              # matched keypoints coincide (same scene)
    # extract_matches runs the pair device-resident (one batched detector / HardNet call); it must return exactly what the
    # one-image-at-a-time composition above gives
    np.testing.assert_array_equal(p1, k1[ids[:, 0]])
    np.testing.assert_array_equal(p2, k2[ids[:, 1]])


def test_config3_pair_vs_reference_and_oracle(detector):
    """BASELINE.json configs[2] against the reference and the CPU oracle at full size (900x1200, seeds 1234 / 1235):
    (i) demo_match.detect against the REFERENCE's own detect() output (tests/golden/r2_detector_large.npz, 2048 of ~2900
    greedy survivors: the top-k cut binds), (ii) the whole pair pipeline -- detector, greedy NMS + sub-pixel, level-1
    patches, HardNet, SMNN -- against oracle.pipeline.extract_matches on the same images (measured recall = precision =
    1.000 with the default precisions, scripts/measure_parity.py)."""
    from balf_b200.configs import config
    from balf_b200.demo import demo_match
    from balf_b200.third_party.hardnet.hardnet_pytorch import HardNet
    from conftest import load_golden
    from oracle import pipeline, weights
    det = copy.deepcopy(detector).to(DEV).eval()
    torch.manual_seed(0)
    hn = HardNet().eval().to(DEV)
    g = load_golden("r2_detector_large.npz")
    args0 = config.default_test_args(sub_pixel=False)
    for seed in (1234, 1235):
        im = synth_u8(900, 1200, seed)
        got = demo_match.detect(args0, im, det, DEV)
        ref = g["detect_900x1200_s%d" % seed]
        assert got.shape == ref.shape == (2048, 3)
        inter = set(map(tuple, got[:, :2])) & set(map(tuple, ref[:, :2]))
        assert len(inter) >= 0.99 * len(ref), (seed, len(inter))
    args = config.default_test_args()
    rgb1 = synth_u8(900, 1200, 1234)
    noise = np.random.default_rng(5).integers(-2, 3, rgb1.shape[:2])[..., None]
    rgb2 = np.clip(rgb1.astype(np.int64) + noise, 0, 255).astype(np.uint8)
    g1, g2 = rgb1[..., 0].copy(), rgb2[..., 0].copy()
    p1, p2 = demo_match.extract_matches(args, rgb1, g1, rgb2, g2, det, hn, DEV)
    w1, w2 = pipeline.extract_matches(args, weights.detector_state_dict(0), weights.hardnet_state_dict(0), rgb1, g1, rgb2, g2)
    got = set(map(tuple, np.round(np.concatenate([p1, p2], 1), 2)))
    want = set(map(tuple, np.round(np.concatenate([w1, w2], 1), 2)))
    assert len(want) > 1500
    assert len(got & want) >= 0.99 * len(want) and len(got & want) >= 0.99 * len(got), (len(got), len(want), len(got & want))
