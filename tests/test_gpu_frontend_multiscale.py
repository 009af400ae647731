"""SURVEY.md section 8(f) rows through the C-ABI against oracle/multiscale.py (parity unpinned: the reference has no
implementation of the multi-scale extraction and Pillow is not importable; see the oracle's header):
RGB -> L conversion, float-image padding, pyramid-level resize (bit-exact), merge of per-level lists (bit-exact,
ties included), the validation extraction ``train_utils.extract_detections`` and the end-to-end multi-scale detect."""
import copy

import numpy as np
import pytest
import torch

from conftest import synth_u8
from oracle import multiscale as oms
from oracle import pipeline, postproc

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_rgb_to_gray_all_paths():
    import balf_b200._capi as capi
    g = torch.Generator().manual_seed(7)
    rgb = torch.randint(0, 256, (3, 37, 53, 3), generator=g, dtype=torch.uint8)
    rgb[0, 0, :8] = torch.tensor([[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 1, 1], [254, 255, 253], [128, 127, 129]],
                                 dtype=torch.uint8)
    got = capi.rgb_to_gray(rgb.to(DEV)).cpu().numpy()
    np.testing.assert_array_equal(got, oms.rgb_to_gray(rgb.numpy()))
    assert got[0, 0, 0] == 0 and got[0, 0, 1] == 255
    np.testing.assert_array_equal(capi.rgb_to_gray(rgb[1].to(DEV)).cpu().numpy(), oms.rgb_to_gray(rgb[1].numpy()))
    with pytest.raises(RuntimeError):
        capi.rgb_to_gray(rgb)


def test_preprocess_f32_matches_reference_padding():
    import balf_b200._capi as capi
    for h, w, c in ((121, 187, 3), (128, 192, 3), (65, 64, 1)):
        img = torch.rand(2, h, w, c, generator=torch.Generator().manual_seed(h))
        x, (top, left) = capi.preprocess_f32(img.to(DEV))
        for b in range(2):
            a = img[b].numpy()
            if c == 1:
                a = np.repeat(a, 3, axis=2)
            want = postproc.mod_padding_symmetric(postproc.make_shape_even(a), 64).transpose(2, 0, 1)
            np.testing.assert_array_equal(x[b].cpu().numpy(), want)


@pytest.mark.parametrize("h,w,c,scale,level", [(480, 640, 1, 0.7, 1), (480, 640, 3, 0.7, 2), (129, 97, 3, 0.5, 1), (200, 300, 1, 2 ** -0.5, 3)])
def test_resize_level_bit_exact(h, w, c, scale, level):
    import balf_b200._capi as capi
    g = torch.Generator().manual_seed(h + level)
    img = torch.randint(0, 256, (2, h, w, c), generator=g, dtype=torch.uint8)
    hs, ws = oms.level_size(h, scale, level), oms.level_size(w, scale, level)
    assert (hs, ws) == (capi.level_size(h, scale, level), capi.level_size(w, scale, level))
    x, (top, left) = capi.resize_preprocess_u8(img.to(DEV), hs, ws)
    x = x.cpu().numpy()
    for b in range(2):
        lvl = oms.resize_level(img[b].numpy(), hs, ws)
        if c == 1:
            lvl = np.repeat(lvl, 3, axis=2)
        want = postproc.mod_padding_symmetric(postproc.make_shape_even(lvl), 64).transpose(2, 0, 1)
        assert x[b].shape == want.shape
        np.testing.assert_array_equal(x[b], want)


def test_merge_levels_bit_exact_with_ties():
    import balf_b200._capi as capi
    rng = np.random.default_rng(5)
    B, K, L = 3, 257, 3
    lists, ref = [], [[] for _ in range(B)]
    scales = [(1.0, 1.0), (640 / 448, 480 / 336), (640 / 314, 480 / 235)]
    for l in range(L):
        xy = rng.integers(0, 600, (B, K, 2)).astype(np.int32)
        sc = np.round(rng.random((B, K)), 2).astype(np.float32)           # two decimals: plenty of ties across levels
        cnt = np.array([K, K - 57 - l, 0 if l == 1 else 5], dtype=np.int32)
        for b in range(B):
            o = np.argsort(-sc[b, :cnt[b]], kind="stable")
            sc[b, :cnt[b]] = sc[b, :cnt[b]][o]
            xy[b, :cnt[b]] = xy[b, :cnt[b]][o]
            ref[b].append((xy[b, :cnt[b]].astype(np.int64), sc[b, :cnt[b]]))
        lists.append((torch.from_numpy(xy).to(DEV), torch.from_numpy(sc).to(DEV), torch.from_numpy(cnt).to(DEV)))
    for k_out in (100, 257, 700):
        xy_o, sc_o, lv_o, cn_o = (t.cpu().numpy() for t in capi.merge_levels_topk(lists, scales, k_out))
        for b in range(B):
            wxy, wsc, wlv = oms.merge_levels(ref[b], scales, k_out)
            n = len(wsc)
            assert cn_o[b] == n
            np.testing.assert_array_equal(sc_o[b, :n], wsc)
            np.testing.assert_array_equal(lv_o[b, :n], wlv)
            np.testing.assert_array_equal(xy_o[b, :n], wxy)
            assert not sc_o[b, n:].any()


def test_extract_detections_matches_oracle(detector, detector_sd):
    from balf_b200.utils import train_utils
    det = copy.deepcopy(detector).to(DEV).eval()
    det.precision = "fp32"
    im = synth_u8(128, 192, 6)
    pts, smap = train_utils.extract_detections(im.astype(np.float32) / np.float32(255), det, DEV, num_points=25)
    want = pipeline.detect_windowed(detector_sd, im, 15, 15, 25)
    assert pts.shape == want.shape and smap.shape == (1, 128, 192)
    np.testing.assert_array_equal(pts[:, :3], want[:, :3])
    np.testing.assert_allclose(pts[:, 3], want[:, 3], rtol=2e-5)
    # the batched call behind the --nms switch
    u8 = torch.from_numpy(im[None, :, :, :1].copy()).to(DEV)
    xy, sc, _, cnt = train_utils.extract_detections_batch(u8, det, nms="apply_nms", num_points=25)
    np.testing.assert_array_equal(xy[0, :int(cnt[0])].cpu().numpy(), want[:, :2].astype(np.int32))
    xy, sc, _, cnt = train_utils.extract_detections_batch(u8, det, nms="nms_fast", num_points=1000, heatmap_confidence_threshold=0.015)
    args = pipeline.default_args(sub_pixel=False, heatmap_confidence_threshold=0.015, num_features=1000)
    wantg = pipeline.detect(args, detector_sd, im)
    assert set(map(tuple, xy[0, :int(cnt[0])].cpu().numpy().tolist())) == set(map(tuple, wantg[:, :2].astype(int).tolist()))
    # box_nms back end: torchvision.ops.nms semantics (oracle restatement pinned to torchvision) on the oracle's own map
    from oracle import postproc
    xy, sc, _, cnt = train_utils.extract_detections_batch(u8, det, nms="box_nms", num_points=1000, heatmap_confidence_threshold=0.015)
    score = postproc.remove_borders(pipeline.score_map(detector_sd, im), 15)
    keep = postproc.box_nms(score, 4, 0.1, 0.015, 1000)
    ys, xs = np.nonzero(keep)
    got = set(map(tuple, xy[0, :int(cnt[0])].cpu().numpy().tolist()))
    want_b = set(zip(xs.tolist(), ys.tolist()))
    assert len(got & want_b) >= 0.99 * len(want_b) and abs(len(got) - len(want_b)) <= 0.01 * len(want_b) + 1
    s_ = sc[0, :int(cnt[0])].cpu().numpy()
    assert (np.diff(s_) <= 0).all()


def test_multiscale_detect_end_to_end(detector, detector_sd):
    """config 5 shape in miniature: 3 levels, scale 0.7, fp32 detector path against the CPU oracle; and same-input
    parity of the per-level extraction + merge on the GPU's own score maps."""
    from balf_b200.configs import config
    from balf_b200.demo import demo_match
    import balf_b200._capi as capi
    det = copy.deepcopy(detector).to(DEV).eval()
    det.precision = "fp32"
    args = config.default_test_args(sub_pixel=False, num_features=300)
    im = synth_u8(256, 320, 11)[:, :, :1].copy()
    u8 = torch.from_numpy(im[None]).to(DEV)
    xy, sc, lv, cnt = demo_match.detect_multiscale_batch_device(args, u8, det, scale=0.7, levels=3)
    n = int(cnt[0])
    assert n == 300 and set(np.unique(lv[0, :n].cpu().numpy())) == {0, 1, 2}
    s = sc[0, :n].cpu().numpy()
    assert (np.diff(s) <= 0).all()
    # (i) same-input parity: oracle extraction + merge on the GPU score maps of every level
    maps = []
    for l in range(3):
        hs, ws = capi.level_size(256, 0.7, l), capi.level_size(320, 0.7, l)
        x, (top, left) = capi.preprocess_u8(u8) if l == 0 else capi.resize_preprocess_u8(u8, hs, ws)
        with torch.inference_mode():
            maps.append(det(x)["prob"][0, top:top + hs, left:left + ws].cpu().numpy())
    wxy, wsc, wlv = oms.detect_multiscale(None, im, 0.7, 3, 300, score_maps=maps)
    np.testing.assert_array_equal(s, wsc)
    np.testing.assert_array_equal(lv[0, :n].cpu().numpy(), wlv)
    np.testing.assert_array_equal(xy[0, :n].cpu().numpy(), wxy)
    # (ii) end to end against the CPU oracle's own detector
    oxy, osc, olv = oms.detect_multiscale(detector_sd, im, 0.7, 3, 300)
    got = set(zip(lv[0, :n].cpu().numpy().tolist(), map(tuple, xy[0, :n].cpu().numpy().tolist())))
    want = set(zip(olv.tolist(), map(tuple, oxy.tolist())))
    assert len(got & want) >= 0.99 * len(want), (len(got & want), len(want))


def test_multiscale_parser_defaults_with_upsampled_level(detector):
    """the pyramid of the reference's own parser (config_hpatches.py:71-76: scale sqrt(2), 5 down-sampled levels and ONE
    up-sampled level = 7 levels, finest first): same-input parity of extraction + merge with the oracle restatement"""
    from balf_b200.configs import config
    from balf_b200.demo import demo_match
    import balf_b200._capi as capi
    det = copy.deepcopy(detector).to(DEV).eval()
    det.precision = "fp32"
    margs = config.default_multiscale_args()
    pyr = config.multiscale_pyramid(margs)
    assert pyr["levels"] == 6 and pyr["upsampled_levels"] == 1 and abs(pyr["scale"] - 2 ** -0.5) < 1e-12
    args = config.default_test_args(sub_pixel=False, num_features=margs.num_points, nms_size=margs.nms_size, border_size=margs.border_size)
    im = synth_u8(384, 448, 13)[:, :, :1].copy()
    u8 = torch.from_numpy(im[None]).to(DEV)
    xy, sc, lv, cnt = demo_match.detect_multiscale_batch_device(args, u8, det, **pyr)
    n = int(cnt[0])
    assert n == margs.num_points and set(np.unique(lv[0, :n].cpu().numpy())) <= set(range(7)) and 0 in lv[0, :n].cpu().numpy()
    maps = []
    for l in range(-1, 6):
        hs, ws = capi.level_size(384, pyr["scale"], l), capi.level_size(448, pyr["scale"], l)
        x, (top, left) = capi.preprocess_u8(u8) if l == 0 else capi.resize_preprocess_u8(u8, hs, ws)
        if l == -1:
            assert (hs, ws) == (543, 634)
        with torch.inference_mode():
            maps.append(det(x)["prob"][0, top:top + hs, left:left + ws].cpu().numpy())
    wxy, wsc, wlv = oms.detect_multiscale(None, im, pyr["scale"], 6, margs.num_points, score_maps=maps, upsampled_levels=1)
    np.testing.assert_array_equal(sc[0, :n].cpu().numpy(), wsc)
    np.testing.assert_array_equal(lv[0, :n].cpu().numpy(), wlv)
    np.testing.assert_array_equal(xy[0, :n].cpu().numpy(), wxy)
    # the up-sampled level's input itself: bit-exact against the restated resize
    x, (top, left) = capi.resize_preprocess_u8(u8, 543, 634)
    want = oms.resize_level(im, 543, 634)
    np.testing.assert_array_equal(x[0, 0, top:top + 543, left:left + 634].cpu().numpy(), want[:, :, 0])


def test_detect_pipeline_equals_detect_batch(detector):
    """the streaming host-buffer API returns exactly what detect_batch returns, in submission order"""
    from balf_b200.configs import config
    from balf_b200.demo import demo_match
    det = copy.deepcopy(detector).to(DEV).eval()
    args = config.default_test_args(sub_pixel=False, num_features=256)
    batches = [torch.from_numpy(np.stack([synth_u8(96, 128, 20 + 3 * i + j)[:, :, :1] for j in range(3)])).pin_memory() for i in range(4)]
    for nms in ("windowed", "greedy"):
        pipe = demo_match.DetectPipeline(args, det, DEV, nms)
        tickets = [pipe.submit(b) for b in batches]
        with pytest.raises(RuntimeError):
            pipe.submit(batches[0])                       # depth = 4 batches in flight: collect first
        for b, t in zip(batches, tickets):
            got = pipe.result(t)
            want = demo_match.detect_batch(args, b, det, DEV, nms)
            for g, w in zip(got, want):
                np.testing.assert_array_equal(g, w)


def test_nccl_gather_c_abi_single_rank():
    """balf_gather_keypoints on an ncclComm_t created through the C-ABI (world size 1 here; bench.py --gpus N runs it over
    NVLink): the records come back unchanged, rank-ordered; pack / unpack kernels agree with balf_b200.sharding."""
    import balf_b200._capi as capi
    from balf_b200 import sharding
    uid = capi.nccl_unique_id()
    assert len(uid) == 128
    comm = capi.nccl_comm_create(uid, 1, 0, DEV)
    try:
        g = torch.Generator().manual_seed(3)
        xy = torch.randint(0, 640, (5, 77, 2), generator=g, dtype=torch.int32).to(DEV)
        sc = torch.rand(5, 77, generator=g).to(DEV)
        cnt = torch.randint(0, 78, (5,), generator=g, dtype=torch.int32).to(DEV)
        xa, sa, ca = capi.gather_keypoints(comm, 1, xy, sc, cnt)
        torch.cuda.synchronize()
        assert torch.equal(xa, xy) and torch.equal(sa, sc) and torch.equal(ca, cnt)
        x2, s2, c2 = sharding.unpack_records(sharding.pack_records(xy, sc, cnt))
        assert torch.equal(x2, xy) and torch.equal(s2, sc) and torch.equal(c2, cnt)
    finally:
        capi.nccl_comm_destroy(comm)
