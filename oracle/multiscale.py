"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy) of the section 8(f) rows next to the hot path.

* ``rgb_to_gray``: PIL ``Image.convert('L')`` as used by the reference's ``load_im`` (demo/demo_match.py:13-19):
  ITU-R 601-2 luma in 16.16 fixed point, L = (19595 R + 38470 G + 7471 B + 32768) >> 16 (Pillow ``Convert.c``;
  PINNED: tests/test_media_fixtures.py checks it against PIL's own output on media/im1.jpg, im2.jpg,
  tests/golden/r2_media.npz).
* ``resize_preprocess`` / ``merge_levels`` / ``detect_multiscale``: the multi-scale pyramid extraction.  The reference
  ships only its argument parser (balf/configs/config_hpatches.py:50-82), no implementation: the semantics are the
  ones documented in include/balf_b200.h, restated here operation by operation in float32.  PARITY UNPINNED.
"""
import numpy as np

from . import detector as odet
from . import postproc

F = np.float32


def rgb_to_gray(rgb):
    r, g, b = (rgb[..., i].astype(np.uint32) for i in range(3))
    return ((19595 * r + 38470 * g + 7471 * b + 32768) >> 16).astype(np.uint8)


def level_size(n, scale, level):
    return max(int(n * (scale ** level) + 0.5), 32)


def _src(n_dst, n_src):
    scale = F(n_src) / F(n_dst)
    s = (np.arange(n_dst, dtype=F) + F(0.5)) * scale - F(0.5)
    s = np.maximum(s, F(0)).astype(F)
    i0 = np.minimum(s.astype(np.int64), n_src - 1)
    i1 = i0 + (i0 < n_src - 1)
    return i0, i1, (s - i0.astype(F)).astype(F)


def resize_level(img, hs, ws):
    """uint8 [H,W,C] -> float32 [hs,ws,C] in [0,1]: bilinear, half-pixel centres, no antialiasing, then /255."""
    H, W = img.shape[:2]
    y0, y1, ly = _src(hs, H)
    x0, x1, lx = _src(ws, W)
    p = img.astype(F)
    hy, hx = (F(1) - ly)[:, None, None], (F(1) - lx)[None, :, None]
    ly, lx = ly[:, None, None], lx[None, :, None]
    a = hx * p[y0][:, x0] + lx * p[y0][:, x1]
    d = hx * p[y1][:, x0] + lx * p[y1][:, x1]
    return ((hy * a + ly * d) / F(255)).astype(F)


def merge_levels(lists, scales, k_out):
    """lists: [(xy int [n,2], score f32 [n])] per level, each ordered; -> (xy f32 [m,2], score [m], level [m])."""
    xs, ss, ls = [], [], []
    for l, ((xy, sc), (sx, sy)) in enumerate(zip(lists, scales)):
        x0 = (xy[:, 0].astype(F) + F(0.5)) * F(sx) - F(0.5)
        y0 = (xy[:, 1].astype(F) + F(0.5)) * F(sy) - F(0.5)
        xs.append(np.stack([x0, y0], 1).astype(F)); ss.append(sc.astype(F)); ls.append(np.full(len(sc), l, np.int32))
    xy, sc, lv = np.concatenate(xs), np.concatenate(ss), np.concatenate(ls)
    order = np.argsort(-sc, kind="stable")[:k_out]          # level-major concatenation + stable sort = the tie rule
    return xy[order], sc[order], lv[order]


def level_score_map(sd, level_f32):
    """padded forward + centre-crop un-pad of one float level image [hs,ws,3] (demo_match.py:22-43 on a float image)."""
    import torch
    hs, ws = level_f32.shape[:2]
    pad = postproc.mod_padding_symmetric(postproc.make_shape_even(level_f32), 64)
    x = torch.from_numpy(np.ascontiguousarray(pad.astype(F).transpose(2, 0, 1))[None])
    with torch.inference_mode():
        prob = odet.detector_forward(sd, x)["prob"][0].numpy()
    _, _, _, _, t0, l0 = postproc.padded_geometry(hs, ws)
    return prob[t0:t0 + hs, l0:l0 + ws]


def detect_multiscale(sd, im_u8, scale, levels, k, nms_size=15, border=15, score_maps=None, upsampled_levels=0):
    """full CPU pipeline for one uint8 image [H,W,C]: per level resize -> pad -> detector -> windowed NMS top-k -> merge.
    ``score_maps`` (optional, one per level) replaces the detector forward (same-input parity of the extraction)."""
    H, W = im_u8.shape[:2]
    lists, scales = [], []
    for i, l in enumerate(range(-int(upsampled_levels), levels)):     # finest (most up-sampled) level first
        hs, ws = level_size(H, scale, l), level_size(W, scale, l)
        if score_maps is not None:
            crop = score_maps[i]
        else:
            lvl = (im_u8.astype(F) / F(255)) if l == 0 else resize_level(im_u8, hs, ws)
            if lvl.shape[2] == 1:
                lvl = np.repeat(lvl, 3, axis=2)
            crop = level_score_map(sd, lvl)
        pts = postproc.windowed_detect(crop, border, nms_size, k)
        lists.append((pts[:, :2].astype(np.int64), pts[:, 3].astype(F)))
        scales.append((W / ws, H / hs))
    return merge_levels(lists, scales, k)
