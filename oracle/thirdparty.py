"""Oracle: restatements of the un-vendored third-party calls on the BALF hot path.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

PARITY UNPINNED.  The reference calls these through two pip dependencies that are
neither vendored under /root/reference nor installable offline:

* ``kornia==0.7.4`` (requirements.txt:3): ``feature.laf_from_center_scale_ori``
  (demo/demo_match.py:65), ``feature.extract_patches_from_pyramid``
  (demo/demo_match.py:69), ``feature.match_smnn`` (demo/demo_match.py:106);
* ``torchgeometry>=0.1.2`` (requirements.txt:7): ``contrib.SpatialSoftArgmax2d``
  (balf/utils/test_utils.py:198-202).

The reference holds no test, fixture or golden vector for any of them, so the
functions below restate the libraries' published algorithms (SURVEY.md appendix
B) and are anchored only on the reference's call sites and argument values.
They are written with torch CPU primitives (``affine_grid``-free closed forms,
``grid_sample``, ``conv2d``, ``interpolate``, ``cdist``, ``topk``) that the two
libraries themselves bottom out in.
"""
import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- torchgeometry
def spatial_soft_argmax2d(logp):
    """torchgeometry.contrib.SpatialSoftArgmax2d(normalized_coordinates=False) on [N,ps,ps].

    e = exp(x - max x); centroid = sum(pos * e) / (sum e + 1e-6); positions 0..ps-1.
    Returns [N,2] = (x, y) float32.  (-inf entries contribute e = 0.)"""
    n, ps, _ = logp.shape
    x = torch.as_tensor(logp, dtype=torch.float32).reshape(n, -1)
    e = torch.exp(x - x.max(dim=-1, keepdim=True)[0])
    inv = 1.0 / (e.sum(-1, keepdim=True) + 1e-6)
    pos = torch.arange(ps, dtype=torch.float32)
    py = pos.view(ps, 1).expand(ps, ps).reshape(-1)
    px = pos.view(1, ps).expand(ps, ps).reshape(-1)
    ex = (e * px).sum(-1, keepdim=True) * inv
    ey = (e * py).sum(-1, keepdim=True) * inv
    return torch.cat([ex, ey], -1).numpy()


# ----------------------------------------------------------------------------- kornia: LAF + pyramid patches
def laf_from_center_scale_ori(xy, scale, angle_deg=None):
    """kornia.feature.laf_from_center_scale_ori: [[s cos, s sin, x], [-s sin, s cos, y]].
    xy [N,2], scale [N] -> [N,2,3].  demo_match.py:62-65 passes scale = s_mult, angle 0."""
    xy = torch.as_tensor(xy, dtype=torch.float32)
    n = xy.shape[0]
    s = torch.as_tensor(scale, dtype=torch.float32).expand(n)
    a = torch.zeros(n) if angle_deg is None else torch.deg2rad(torch.as_tensor(angle_deg, dtype=torch.float32))
    c, sn = torch.cos(a), torch.sin(a)
    rot = torch.stack([torch.stack([c, sn], -1), torch.stack([-sn, c], -1)], -2)
    return torch.cat([rot * s.view(n, 1, 1), xy.view(n, 2, 1)], -1)


_K5 = torch.tensor([1.0, 4.0, 6.0, 4.0, 1.0])


def pyrdown(img):
    """kornia.geometry.transform.pyrdown: 5x5 binomial blur (reflect border) then bilinear
    resize (align_corners=False) to (int(h/2), int(w//2)).  img [1,1,h,w] fp32."""
    h, w = img.shape[-2:]
    k = (_K5.view(5, 1) * _K5.view(1, 5) / 256.0).view(1, 1, 5, 5)
    blur = F.conv2d(F.pad(img, (2, 2, 2, 2), mode="reflect"), k)
    return F.interpolate(blur, size=(int(float(h) / 2.0), int(float(w) // 2.0)), mode="bilinear",
                         align_corners=False)


def extract_patches_from_pyramid(img, laf, PS=32):
    """kornia.feature.extract_patches_from_pyramid(img[1,1,H,W], laf[N,2,3], PS) with
    normalize_lafs_before_extraction=True -> [N,1,PS,PS].  Also returns the per-keypoint
    pyramid level (all 1 for the demo's s_mult=60, PS=32)."""
    _, _, H, W = img.shape
    n = laf.shape[0]
    m0 = float(min(H - 1, W - 1))
    A = laf[:, :, :2] / m0                                    # normalize_laf
    cx, cy = laf[:, 0, 2] / float(W - 1), laf[:, 1, 2] / float(H - 1)
    det = (A[:, 0, 0] * A[:, 1, 1] - A[:, 1, 0] * A[:, 0, 1]) * m0 * m0
    scale = 2.0 * torch.sqrt(det.abs() + 1e-10) / float(PS)
    max_level = min(H, W) // PS
    level = torch.log2(scale).clamp(min=0.0, max=float(max(0, max_level - 1))).long()
    out = torch.zeros(n, 1, PS, PS)
    cur, lvl = img, 0
    u = (2.0 * torch.arange(PS, dtype=torch.float32) + 1.0) / PS - 1.0   # affine_grid base, align_corners=False
    while min(cur.shape[-2:]) >= PS:
        h, w = cur.shape[-2:]
        sel = torch.nonzero(level == lvl).flatten()
        if sel.numel():
            ml = float(min(h - 1, w - 1))
            a = A[sel] * ml                                   # denormalize_laf at this level
            tx, ty = cx[sel] * (w - 1), cy[sel] * (h - 1)
            gx = a[:, 0, 0, None, None] * u.view(1, 1, PS) + a[:, 0, 1, None, None] * u.view(1, PS, 1) + tx[:, None, None]
            gy = a[:, 1, 0, None, None] * u.view(1, 1, PS) + a[:, 1, 1, None, None] * u.view(1, PS, 1) + ty[:, None, None]
            grid = torch.stack([2.0 * gx / float(w - 1) - 1.0, 2.0 * gy / float(h - 1) - 1.0], -1)
            # one grid_sample over a [1,1,h,w] image with the keypoints stacked along the grid's
            # row axis -- identical arithmetic to kornia's expand()-ed batch.
            pat = F.grid_sample(cur, grid.reshape(1, -1, PS, 2), padding_mode="border", align_corners=False)
            out[sel] = pat.reshape(len(sel), 1, PS, PS)
        cur = pyrdown(cur)
        lvl += 1
    return out, level


# ----------------------------------------------------------------------------- kornia: SMNN matching
def distance_matrix(d1, d2):
    """torch.cdist(d1, d2) -- the matmul form sqrt(clamp(|a|^2 + |b|^2 - 2ab, 0)) that ATen
    uses for > 25 rows."""
    return torch.cdist(torch.as_tensor(d1), torch.as_tensor(d2), compute_mode="use_mm_for_euclid_dist")


def _snn(dm, th):
    if dm.shape[1] < 2:
        return torch.zeros(0), torch.zeros(0, 2, dtype=torch.long)
    # two smallest per row; ties -> lowest column first (canonical; torch.topk is unspecified)
    order = torch.sort(dm, dim=1, stable=True)
    v0, v1 = order.values[:, 0], order.values[:, 1]
    ratio = v0 / v1
    keep = ratio <= th
    rows = torch.nonzero(keep).flatten()
    return ratio[keep], torch.stack([rows, order.indices[:, 0][keep]], 1)


def match_smnn(d1, d2, th=0.99, dm=None):
    """kornia.feature.match_smnn: SNN ratio test in both directions, keep mutual pairs, sort
    by the first index, distance = max of the two ratios.  Returns (dists[M,1], idxs[M,2])."""
    d1, d2 = torch.as_tensor(d1), torch.as_tensor(d2)
    if d1.shape[0] < 2 or d2.shape[0] < 2:
        return torch.zeros(0, 1), torch.zeros(0, 2, dtype=torch.long)
    if dm is None:
        dm = distance_matrix(d1, d2)
    r12, m12 = _snn(dm, th)
    r21, m21 = _snn(dm.t(), th)
    if len(r12) == 0 or len(r21) == 0:
        return torch.zeros(0, 1), torch.zeros(0, 2, dtype=torch.long)
    back = torch.full((d2.shape[0],), -1, dtype=torch.long)
    back[m21[:, 0]] = m21[:, 1]
    rback = torch.zeros(d2.shape[0], dtype=dm.dtype)
    rback[m21[:, 0]] = r21
    mutual = back[m12[:, 1]] == m12[:, 0]
    idx = m12[mutual]
    dist = torch.maximum(r12[mutual], rback[idx[:, 1]])
    return dist.view(-1, 1), idx
