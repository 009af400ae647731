"""Drop-in for the validation-time extraction of the reference's ``balf/utils/train_utils.py``.

``extract_detections`` (train_utils.py:416-453) keeps its signature and return layout; the batched
``extract_detections_batch`` puts the NMS back ends the evaluation parsers advertise
(``--nms {apply_nms, apply_nms_fast, nms_fast, box_nms}``, balf/configs/config_hpatches.py:25-26) behind one call.
Everything runs through the C-ABI kernels (no CPU path)."""
import numpy as np
import torch

from .. import _capi

# back ends of the reference's --nms switch -> kernel family
#   apply_nms       repeatability_tools.py:19-23  windowed maximum filter + k-th value top-k      -> windowed
#   nms_fast        repeatability_tools.py:138+   greedy, candidates >= threshold                 -> greedy
#   apply_nms_fast  repeatability_tools.py:25+    greedy over every pixel of the map; the validation loop
#                   (train_utils.py:165-170) then keeps the points >= 0.015                       -> greedy
#   box_nms         repeatability_tools.py:227+   torchvision.ops.nms on 4x4 boxes (IoU 0.1), then the k best  -> box
NMS_BACKENDS = {"apply_nms": "windowed", "nms_fast": "greedy", "apply_nms_fast": "greedy", "box_nms": "box"}


def _points(xy, sc, n):
    pts = np.zeros((n, 4), dtype=np.float64)
    pts[:, 0:2] = xy[:n]
    pts[:, 2] = 1.0
    pts[:, 3] = sc[:n]
    return pts


@torch.no_grad()
def extract_detections(image_RGB_norm, model, device, cell_size=8, nms_size=15, num_points=25, border_size=15):
    """image_RGB_norm [H,W,3] float in [0,1] -> (pts [<=num_points, 4] float64 rows (x, y, 1.0, score) ordered by
    score, score_map_batch [1,H,W] = the un-padded score map on ``device``).
    (The reference indexes the 3-D ``prob`` with four indices at train_utils.py:441, which raises; the crop intended
    there is returned.)"""
    if cell_size != 8:
        raise ValueError("only cell_size = 8 is built")
    dev = torch.device(device)
    img = torch.as_tensor(np.ascontiguousarray(image_RGB_norm), dtype=torch.float32).to(dev)[None]
    h, w = img.shape[1], img.shape[2]
    x, (top, left) = _capi.preprocess_f32(img)
    prob = model(x, precision=model.resolve_precision("windowed"), want_logits=False)["prob"]
    xy, sc, cnt = _capi.windowed_nms_topk(prob, num_points, border=border_size, nms_size=nms_size, crop=(top, left, h, w))
    n = int(cnt[0])
    return _points(xy[0].cpu().numpy(), sc[0].cpu().numpy(), n), prob[:, top:top + h, left:left + w]


@torch.no_grad()
def extract_detections_batch(images_u8, model, nms="nms_fast", nms_size=15, num_points=1000, border_size=15,
                             heatmap_confidence_threshold=0.015, sub_pixel=False, patch_size=5, box_size=4, box_iou=0.1):
    """Batched evaluation-time extraction with the reference's evaluation defaults (config_hpatches.py:25-44).
    images_u8 [B,H,W,C] uint8 CUDA -> (xy int32 [B,K,2], score fp32 [B,K], dxdy fp32 [B,K,2] | None, count int32 [B])
    on the device."""
    if nms not in NMS_BACKENDS:
        raise ValueError("nms must be one of %s" % sorted(NMS_BACKENDS))
    _, h, w, _ = images_u8.shape
    x, (top, left) = _capi.preprocess_u8(images_u8)
    kind = NMS_BACKENDS[nms]
    prob = model(x, precision=model.resolve_precision("windowed" if kind == "windowed" else "greedy"), want_logits=False)["prob"]
    if kind == "box":
        # repeatability_tools.box_nms(prob, size=4, iou=0.1, min_prob, keep_top_k) on the border-masked crop, then the
        # surviving pixels as points (get_point_coordinates): a 1 x 1 "window" keeps every positive pixel
        crop = _capi.apply_nms_map(prob[:, top:top + h, left:left + w].contiguous(), 1, border=border_size)
        kept = _capi.box_nms_map(crop, size=box_size, iou=box_iou, min_prob=heatmap_confidence_threshold, keep_top_k=num_points)
        xy, sc, cnt = _capi.windowed_nms_topk(kept, num_points, border=0, nms_size=1)
        return xy, sc, None, cnt
    if kind == "windowed":
        xy, sc, cnt = _capi.windowed_nms_topk(prob, num_points, border=border_size, nms_size=nms_size, crop=(top, left, h, w))
        return xy, sc, None, cnt
    return _capi.greedy_nms_topk(prob, num_points, border=border_size, thr=heatmap_confidence_threshold, radius=nms_size,
                                 subpixel_ps=patch_size if sub_pixel else 0, crop=(top, left, h, w))
