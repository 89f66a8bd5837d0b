"""CPU restatement (numpy, fp32) of the reference's result-side helpers that follow NMS.

TEST INFRASTRUCTURE ONLY - never imported by the product package (see oracle/__init__.py).

Paths are relative to /root/reference/ultralytics/.  Unlike the reference (in-place torch ops) these are pure functions
on numpy float32 arrays; every arithmetic step is one IEEE fp32 operation, in the reference's order, so results equal the
reference's CPU results bit for bit.  Pinned by tests/golden/post/result_ops.npz (oracle/make_golden_post.py runs the live
reference).
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def letterbox_scalars(img1_shape, img0_shape, ratio_pad=None):
    """gain, box pad (x, y), coord pad (x, y): utils/ops.py:120-127 and :580-587."""
    h0, w0 = img0_shape[:2]
    if ratio_pad is None:
        h1, w1 = img1_shape[:2]
        gain = min(h1 / h0, w1 / w0)
        cpad = ((w1 - w0 * gain) / 2, (h1 - h0 * gain) / 2)
        pad = (round(cpad[0] - 0.1), round(cpad[1] - 0.1))
    else:
        gain = ratio_pad[0][0]
        pad = cpad = tuple(ratio_pad[1])
    return gain, pad, cpad


def _clamp(v, hi):
    with np.errstate(invalid="ignore"):
        out = np.minimum(np.maximum(v, F(0)), F(hi))
    return np.where(np.isnan(v), v, out).astype(F)


def clip_boxes_oracle(boxes, shape):
    """utils/ops.py:152-177: x to [0, w], y to [0, h]."""
    h, w = shape[:2]
    b = np.array(boxes, dtype=F, copy=True)
    for col, hi in ((0, w), (1, h), (2, w), (3, h)):
        b[..., col] = _clamp(b[..., col], hi)
    return b


def scale_boxes_oracle(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True, xywh=False):
    """utils/ops.py:102-135."""
    gain, pad, _ = letterbox_scalars(img1_shape, img0_shape, ratio_pad)
    b = np.array(boxes, dtype=F, copy=True)
    if padding:
        for col in (0, 1) if xywh else (0, 1, 2, 3):
            b[..., col] = b[..., col] - F(pad[col & 1])
    b[..., :4] = b[..., :4] / F(gain)
    return b if xywh else clip_boxes_oracle(b, img0_shape)


def _remainder(a, m):
    """torch.remainder for fp32 (fmod, then moved into the divisor's sign); fmod is exact."""
    r = np.fmod(a, F(m)).astype(F)
    fix = (r != 0) & ((F(m) < 0) != (r < 0))
    return np.where(fix, r + F(m), r).astype(F)


def regularize_rboxes_oracle(rboxes):
    """utils/ops.py:621-636."""
    r = np.array(rboxes, dtype=F, copy=True)
    t = r[..., 4]
    swap = _remainder(t, math.pi) >= F(math.pi / 2)
    w, h = r[..., 2].copy(), r[..., 3].copy()
    r[..., 2] = np.where(swap, h, w)
    r[..., 3] = np.where(swap, w, h)
    r[..., 4] = _remainder(t, math.pi / 2)
    return r


def clip_coords_oracle(coords, shape):
    """utils/ops.py:598-618."""
    h, w = shape[:2]
    c = np.array(coords, dtype=F, copy=True)
    c[..., 0] = _clamp(c[..., 0], w)
    c[..., 1] = _clamp(c[..., 1], h)
    return c


def scale_coords_oracle(img1_shape, coords, img0_shape, ratio_pad=None, normalize=False, padding=True):
    """utils/ops.py:562-595."""
    gain, _, cpad = letterbox_scalars(img1_shape, img0_shape, ratio_pad)
    h0, w0 = img0_shape[:2]
    c = np.array(coords, dtype=F, copy=True)
    if padding:
        c[..., 0] = c[..., 0] - F(cpad[0])
        c[..., 1] = c[..., 1] - F(cpad[1])
    c[..., 0] = c[..., 0] / F(gain)
    c[..., 1] = c[..., 1] / F(gain)
    c = clip_coords_oracle(c, img0_shape)
    if normalize:
        c[..., 0] = c[..., 0] / F(w0)
        c[..., 1] = c[..., 1] / F(h0)
    return c


def obb_result_oracle(pred, img1_shape, img0_shape):
    """models/yolo/obb/predict.py:59-61: rows cx,cy,w,h,conf,cls,angle -> (N, 7) x,y,w,h,angle,conf,cls."""
    p = np.asarray(pred, dtype=F)
    rb = regularize_rboxes_oracle(np.concatenate([p[:, :4], p[:, -1:]], -1))
    rb[:, :4] = scale_boxes_oracle(img1_shape, rb[:, :4], img0_shape, xywh=True)
    return np.concatenate([rb, p[:, 4:6]], -1)


def kpts_decode_oracle(kpts, level_hw, strides, kpt_shape):
    """nn/modules/head.py:1254-1273 (Pose.kpts_decode, non-export branch) on a torch tensor (B, nk*ndim, A) of any float
    dtype; every operation runs in that dtype like the reference's.  Anchors/strides as cached by head.py:163-165."""
    import torch

    from oracle.postproc_oracle import anchor_table

    nk, ndim = kpt_shape
    b, _, a = kpts.shape
    anchors, srow = anchor_table(level_hw, strides, kpts.dtype)  # (2, A), (1, A)
    y = kpts.reshape(b, nk, ndim, a)
    parts = [(y[:, :, 0] * 2.0 + (anchors[0] - 0.5)) * srow, (y[:, :, 1] * 2.0 + (anchors[1] - 0.5)) * srow]
    if ndim == 3:
        parts.append(y[:, :, 2].sigmoid())
    for d in range(len(parts), ndim):
        parts.append(y[:, :, d])
    return torch.stack(parts, 2).reshape(b, nk * ndim, a)


def crop_mask_oracle(masks, boxes):
    """utils/ops.py:464-486, the comparison branch (taken for CUDA tensors and for n >= 50): pixel (row c, col r) survives
    iff x1 <= r < x2 and y1 <= c < y2.  (For n < 50 on the CPU the reference instead slices at ``boxes.round().int()``,
    ops.py:476-481 - a different rounding that the CUDA path of the reference never takes.)"""
    import torch

    n, h, w = masks.shape
    x1, y1, x2, y2 = (boxes[:, i].reshape(n, 1, 1) for i in range(4))
    cols = torch.arange(w, dtype=boxes.dtype).reshape(1, 1, w)
    rows = torch.arange(h, dtype=boxes.dtype).reshape(1, h, 1)
    inside = (cols >= x1) & (cols < x2) & (rows >= y1) & (rows < y2)
    return masks * inside


def process_mask_oracle(protos, masks_in, bboxes, shape, upsample=False):
    """utils/ops.py:489-513.  Returns (uint8 masks, the float values that were thresholded)."""
    import torch
    import torch.nn.functional as F

    c, mh, mw = protos.shape
    low = (masks_in @ protos.float().reshape(c, mh * mw)).reshape(-1, mh, mw)
    rw, rh = mw / shape[1], mh / shape[0]
    low = crop_mask_oracle(low, bboxes * torch.tensor([[rw, rh, rw, rh]]))
    if upsample:
        low = F.interpolate(low[None], tuple(shape), mode="bilinear")[0]
    return (low > 0).to(torch.uint8), low


def scale_masks_window(mh, mw, shape, padding=True):
    """utils/ops.py:544-559: the un-padded window of the mask grid."""
    gain = min(mh / shape[0], mw / shape[1])
    pad_w, pad_h = mw - shape[1] * gain, mh - shape[0] * gain
    if padding:
        pad_w, pad_h = pad_w / 2, pad_h / 2
    top, left = (round(pad_h - 0.1), round(pad_w - 0.1)) if padding else (0, 0)
    return top, left, mh - round(pad_h + 0.1), mw - round(pad_w + 0.1)


def process_mask_native_oracle(protos, masks_in, bboxes, shape):
    """utils/ops.py:516-541."""
    import torch
    import torch.nn.functional as F

    c, mh, mw = protos.shape
    low = (masks_in @ protos.float().reshape(c, mh * mw)).reshape(-1, mh, mw)
    top, left, bottom, right = scale_masks_window(mh, mw, shape)
    up = F.interpolate(low[None, :, top:bottom, left:right], tuple(shape), mode="bilinear")[0]
    up = crop_mask_oracle(up, bboxes)
    return (up > 0).to(torch.uint8), up


def box_iou_oracle(box1, box2, eps=1e-7):
    """utils/metrics.py:54-75 on numpy fp32: (N, 4) x (M, 4) xyxy -> (N, M)."""
    a = np.asarray(box1, dtype=F)[:, None, :]
    b = np.asarray(box2, dtype=F)[None, :, :]
    wh = np.maximum(np.minimum(a[..., 2:], b[..., 2:]) - np.maximum(a[..., :2], b[..., :2]), F(0))
    inter = wh[..., 0] * wh[..., 1]
    area_a = (a[..., 2] - a[..., 0]) * (a[..., 3] - a[..., 1])
    area_b = (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    return (inter / (area_a + area_b - inter + F(eps))).astype(F)


def match_predictions_oracle(pred_classes, true_classes, iou, iouv):
    """engine/validator.py:267-307, the non-scipy branch, restated without the sort: per IoU level a detection is correct
    iff (a) its best same-class label reaches the level and (b) it is the lowest-index detection among those sharing that
    best label at this level.  (The reference sorts the candidate pairs by IoU, keeps the first pair of every detection and
    then - the pairs now being in detection order - the first pair of every label.)  Scores must be tie-free."""
    pc = np.asarray(pred_classes, dtype=F)
    tc = np.asarray(true_classes, dtype=F)
    m = np.asarray(iou, dtype=F) * (tc[:, None] == pc[None, :]).astype(F)  # (labels, detections)
    n = pc.shape[0]
    correct = np.zeros((n, len(iouv)), dtype=bool)
    if m.size == 0:
        return correct
    best_label = m.argmax(0)
    best = m.max(0)
    for i, t in enumerate(iouv):
        taken = set()
        for d in range(n):
            if best[d] > 0 and best[d] >= F(t) and best_label[d] not in taken:
                taken.add(best_label[d])
                correct[d, i] = True
    return correct


def nms_model_oracle(pred, image_hw, nc, conf, iou, max_det, agnostic_nms=False):
    """engine/exporter.py:1417-1481 (NMSModel.forward after self.model(x); detect / segment / pose, non-TF formats): torch CPU,
    torchvision.ops.nms on the normalised + class-offset boxes.  pred (B, 4+nc+extra, A), boxes xyxy.  -> (B, max_det', 6+extra)."""
    import torch
    from torchvision.ops import nms

    p = pred.transpose(-1, -2)
    extra = p.shape[-1] - 4 - nc
    boxes, scores, extras = p.split([4, nc, extra], dim=2)
    scores, classes = scores.max(dim=-1)
    max_det = min(p.shape[1], max_det)
    out = torch.zeros(p.shape[0], max_det, 6 + extra, dtype=p.dtype)
    mult = 1 / max(nc, 1)
    side = torch.tensor(tuple(image_hw), dtype=p.dtype).max()
    for i in range(p.shape[0]):
        keep_mask = scores[i] > conf
        box, score, cls, ext = boxes[i][keep_mask], scores[i][keep_mask], classes[i][keep_mask], extras[i][keep_mask]
        nb = mult * (box / side)
        if not agnostic_nms:
            nb = nb + (cls.view(-1, 1) * mult)
        keep = nms(nb, score, iou)[:max_det]
        dets = torch.cat([box[keep], score[keep].view(-1, 1), cls[keep].view(-1, 1).to(out.dtype), ext[keep]], dim=-1)
        out[i, : dets.shape[0]] = dets
    return out


def detect_postprocess_oracle(preds, max_det, nc):
    """nn/modules/head.py:193-214 (Detect.postprocess) restated as ONE global ranking: the K = min(max_det, A) best
    (anchor, class) pairs by score (stable: lower flat index first), rows box(4), score, class.  Equal to the reference's
    two-stage top-k whenever scores are tie-free (see ultralytics_pro_b200.head.detect_postprocess)."""
    import torch

    b, a, _ = preds.shape
    k = min(max_det, a)
    flat = preds[..., 4:].reshape(b, a * nc)
    order = torch.sort(flat, dim=1, descending=True, stable=True).indices[:, :k]
    rows = torch.arange(b)[:, None]
    return torch.cat([preds[rows, order // nc, :4], flat[rows, order][..., None], (order % nc)[..., None].to(preds.dtype)], -1)
