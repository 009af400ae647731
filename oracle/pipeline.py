"""Oracle: the demo pipeline (detect -> extract_features -> extract_matches) on CPU.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Follows
``/root/reference/demo/demo_match.py:21-112`` and
``/root/reference/balf/utils/train_utils.py:416-453`` (windowed variant).
Pieces that go through kornia / torchgeometry are PARITY UNPINNED (``thirdparty.py``).
"""
import types

import numpy as np
import torch

from . import detector, hardnet, postproc, thirdparty


def default_args(**kw):
    """CLI defaults of balf/configs/config.py:35-65 (parse_test_config)."""
    a = dict(border_size=15, nms_size=15, num_features=2048, s_mult=60, order_coord="xysr",
             heatmap_confidence_threshold=0.001, sub_pixel=True, patch_size=4)
    a.update(kw)
    return types.SimpleNamespace(**a)


def score_map(sd, im_u8):
    """demo_match.py:22-43 -- pad, forward, centre-crop un-pad.  im_u8: [H,W,3] uint8."""
    h, w = im_u8.shape[:2]
    x = torch.from_numpy(postproc.preprocess(im_u8))
    with torch.inference_mode():
        prob = detector.detector_forward(sd, x)["prob"][0].numpy()
    _, _, _, _, hs, ws = postproc.padded_geometry(h, w)
    return prob[hs:hs + h, ws:ws + w]


def detect_from_score_map(args, score, nms=postproc.greedy_nms):
    """demo_match.py:44-57 -- border mask, threshold + greedy NMS (+ sub-pixel), top-k.
    Returns ([K,3] (x, y, 1.0) float64, [K] scores)."""
    pts = postproc.get_points_direct_from_score_map(
        postproc.remove_borders(score, args.border_size), args.heatmap_confidence_threshold,
        args.nms_size, args.sub_pixel, args.patch_size, order_coord=args.order_coord, nms=nms)
    if pts.size == 0:
        return np.zeros((0, 3)), np.zeros((0,))
    top = pts[np.argsort(-pts[:, 3], kind="stable")][:args.num_features]
    return top[:, 0:3], top[:, 3]


def detect(args, sd, im_u8, nms=postproc.greedy_nms):
    return detect_from_score_map(args, score_map(sd, im_u8), nms)[0]


def detect_windowed(sd, im_u8, border=15, nms_size=15, num_points=2048):
    """train_utils.py:416-453 (extract_detections): windowed NMS + k-th-value top-k."""
    return postproc.windowed_detect(score_map(sd, im_u8), border, nms_size, num_points)


def describe(args, hn_sd, im_gray_u8, kpts_xy, chunk=1000):
    """demo_match.py:62-93 -- LAF (scale s_mult), level-1 pyramid patches, HardNet in chunks."""
    kp = torch.as_tensor(np.asarray(kpts_xy, np.float64)).float()
    laf = thirdparty.laf_from_center_scale_ori(kp, float(args.s_mult))
    img = torch.from_numpy(im_gray_u8)[None, None].float() / 255.0
    patches, _ = thirdparty.extract_patches_from_pyramid(img, laf, 32)
    with torch.inference_mode():
        descs = [hardnet.hardnet_forward(hn_sd, patches[i:i + chunk]) for i in range(0, len(patches), chunk)]
    return (torch.cat(descs) if descs else torch.zeros(0, 128)).numpy(), patches.numpy()


def extract_features(args, sd, hn_sd, im_rgb, im_gray):
    k = detect(args, sd, im_rgb)
    return k[:, 0:2], describe(args, hn_sd, im_gray, k[:, 0:2])[0]


def extract_matches(args, sd, hn_sd, rgb1, gray1, rgb2, gray2):
    """demo_match.py:97-112."""
    k1, d1 = extract_features(args, sd, hn_sd, rgb1, gray1)
    k2, d2 = extract_features(args, sd, hn_sd, rgb2, gray2)
    _, ids = thirdparty.match_smnn(d1, d2, 0.99)
    ids = ids.numpy()
    return k1[ids[:, 0], :2], k2[ids[:, 1], :2]
