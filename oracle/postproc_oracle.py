"""CPU restatement of the reference's detection post-processing path.

TEST INFRASTRUCTURE ONLY - never imported by the product package (see oracle/__init__.py).

Every function states the reference lines (relative to /root/reference/ultralytics/) whose
arithmetic it restates.  The restatement deliberately uses the same torch CPU operators, in
the same order, on the same dtypes as the reference so that (a) fp32 results are bit-identical
to the reference's and (b) its run time is representative when bench.py times it as the
``cpu_baseline`` ("port").  ``greedy_nms_plain`` / ``fast_nms_plain`` are independent numpy
restatements used to cross-check the third-party ``torchvision.ops.nms`` call.

Pinned by tests/golden/*.npz (generated from the live reference by oracle/make_golden.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

EPS_PROBIOU = 1e-7


# ----------------------------------------------------------------------------------------------
# decode  (nn/modules/head.py:151-191, :1026-1042; nn/modules/block.py:250-253; utils/tal.py:352-403)
# ----------------------------------------------------------------------------------------------
def anchor_table(level_hw, strides, dtype=torch.float32):
    """Grid-centre anchors, level-major then row-major (utils/tal.py:352-364).

    Returns ``(anchors (2, A), stride_row (1, A))`` exactly as cached by head.py:163-165.
    """
    pts, srow = [], []
    for (h, w), s in zip(level_hw, strides):
        gx = torch.arange(w, dtype=dtype) + 0.5
        gy = torch.arange(h, dtype=dtype) + 0.5
        yy, xx = torch.meshgrid(gy, gx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).reshape(-1, 2))
        srow.append(torch.full((h * w, 1), float(s), dtype=dtype))
    return torch.cat(pts).t(), torch.cat(srow).t()


def dfl_expect(box_logits: torch.Tensor, reg_max: int = 16) -> torch.Tensor:
    """Softmax over the ``reg_max`` bins of each side, then the arange-weighted sum.

    block.py:250-253: channel = side*reg_max + bin; the 1x1 conv has frozen weights arange(reg_max)
    (block.py:245-247), i.e. an expectation.  Same operators (softmax on the transposed view, conv2d)
    so the CPU cost and rounding match.
    """
    b, _, a = box_logits.shape
    w = torch.arange(reg_max, dtype=torch.float32).view(1, reg_max, 1, 1).to(box_logits.dtype)
    p = box_logits.view(b, 4, reg_max, a).transpose(2, 1).softmax(1)
    return F.conv2d(p, w).view(b, 4, a)


def decode_oracle(levels, strides, nc: int, reg_max: int = 16, angle: torch.Tensor | None = None,
                  xyxy: bool = False) -> torch.Tensor:
    """Restatement of ``Detect._inference`` (head.py:151-169) / ``OBB`` variant (head.py:1040-1042).

    levels : list of (B, 4*reg_max+nc, H_i, W_i) tensors (any float dtype; arithmetic runs in that dtype,
             as it does in the reference).
    angle  : optional (B, 1, A) *activated* angle ``(sigmoid(t)-0.25)*pi`` (head.py:1031); selects the
             rotated decode ``dist2rbox`` (tal.py:385-403).  The caller concatenates it afterwards
             (head.py:1038), see ``obb_forward_oracle``.
    returns (B, 4+nc, A) in the input dtype: rows cx,cy,w,h (input pixels) then sigmoid scores.
    """
    b = levels[0].shape[0]
    no = 4 * reg_max + nc
    flat = torch.cat([lv.reshape(b, no, -1) for lv in levels], 2)  # head.py:162
    anchors, srow = anchor_table([lv.shape[2:] for lv in levels], strides, flat.dtype)
    box, cls = flat.split((4 * reg_max, nc), 1)  # head.py:167
    dist = dfl_expect(box, reg_max) if reg_max > 1 else box
    anc = anchors.unsqueeze(0)
    if angle is None:
        lt, rb = dist.chunk(2, 1)  # tal.py:369
        p1 = anc - lt
        p2 = anc + rb
        if xyxy:
            dbox = torch.cat((p1, p2), 1)
        else:
            dbox = torch.cat(((p1 + p2) / 2, p2 - p1), 1)  # tal.py:373-375
    else:
        lt, rb = dist.split(2, dim=1)  # tal.py:397
        co, si = torch.cos(angle), torch.sin(angle)
        xf, yf = ((rb - lt) / 2).split(1, dim=1)
        rx, ry = xf * co - yf * si, xf * si + yf * co
        dbox = torch.cat([torch.cat([rx, ry], 1) + anc, lt + rb], 1)  # tal.py:401-403
    return torch.cat((dbox * srow, cls.sigmoid()), 1)  # head.py:168-169


def obb_forward_oracle(levels, angle_logits, strides, nc: int, reg_max: int = 16) -> torch.Tensor:
    """OBB.forward inference branch (head.py:1026-1038): activate angle, rotated decode, append angle."""
    ang = (angle_logits.sigmoid() - 0.25) * math.pi
    y = decode_oracle(levels, strides, nc, reg_max, angle=ang)
    return torch.cat([y, ang], 1)


# ----------------------------------------------------------------------------------------------
# IoU kernels  (utils/metrics.py:54-75, 187-203, 251-284)
# ----------------------------------------------------------------------------------------------
def _cov_terms(b: torch.Tensor):
    """metrics.py:187-203 - Gaussian covariance terms of xywhr boxes."""
    g = torch.cat((b[:, 2:4].pow(2) / 12, b[:, 4:]), dim=-1)
    a, bb, c = g.split(1, dim=-1)
    co, si = c.cos(), c.sin()
    c2, s2 = co.pow(2), si.pow(2)
    return a * c2 + bb * s2, a * s2 + bb * c2, (a - bb) * co * si


def probiou_matrix(o1: torch.Tensor, o2: torch.Tensor, eps: float = EPS_PROBIOU) -> torch.Tensor:
    """metrics.py:251-284 - pairwise ProbIoU of (N,5) and (M,5) xywhr boxes -> (N, M)."""
    x1, y1 = o1[..., :2].split(1, dim=-1)
    x2, y2 = (t.squeeze(-1)[None] for t in o2[..., :2].split(1, dim=-1))
    a1, b1, c1 = _cov_terms(o1)
    a2, b2, c2 = (t.squeeze(-1)[None] for t in _cov_terms(o2))
    den = (a1 + a2) * (b1 + b2) - (c1 + c2).pow(2)
    t1 = (((a1 + a2) * (y1 - y2).pow(2) + (b1 + b2) * (x1 - x2).pow(2)) / (den + eps)) * 0.25
    t2 = (((c1 + c2) * (x2 - x1) * (y1 - y2)) / (den + eps)) * 0.5
    t3 = (den / (4 * ((a1 * b1 - c1.pow(2)).clamp_(0) * (a2 * b2 - c2.pow(2)).clamp_(0)).sqrt() + eps) + eps).log() * 0.5
    bd = (t1 + t2 + t3).clamp(eps, 100.0)
    hd = (1.0 - (-bd).exp() + eps).sqrt()
    return 1 - hd


def box_iou_matrix(b1: torch.Tensor, b2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """metrics.py:54-75 - pairwise IoU with eps in the denominator."""
    (a1, a2), (c1, c2) = b1.float().unsqueeze(1).chunk(2, 2), b2.float().unsqueeze(0).chunk(2, 2)
    inter = (torch.min(a2, c2) - torch.max(a1, c1)).clamp_(0).prod(2)
    return inter / ((a2 - a1).prod(2) + (c2 - c1).prod(2) - inter + eps)


# ----------------------------------------------------------------------------------------------
# suppression  (utils/nms.py:187-296 and torchvision.ops.nms, third-party, unpinned by the reference,
# 0.26.0 in this image - SURVEY.md section 8c)
# ----------------------------------------------------------------------------------------------
def stable_desc_order(scores: torch.Tensor) -> torch.Tensor:
    """Descending order, ties -> lower index first (torchvision's sort; our definition for nms.py:138,217,264)."""
    return torch.sort(scores, descending=True, stable=True).indices


def greedy_nms_plain(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    """Independent numpy restatement of greedy NMS (nms.py:239-296 == torchvision.ops.nms on CPU).

    fp32 arithmetic with every op separately rounded, ``inter / (area_i + area_j - inter)`` with no eps,
    suppress iff that fp32 quotient, widened to double, is ``> thr`` (torchvision's CPU kernel compares the
    scalar_t quotient with the double threshold).  NaN (0/0) never suppresses.
    """
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), np.int64)
    order = np.argsort(-scores.astype(np.float32), kind="stable")
    x1, y1, x2, y2 = (boxes[order, k] for k in range(4))
    area = (x2 - x1) * (y2 - y1)
    dead = np.zeros(n, bool)
    keep = []
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(n):
            if dead[i]:
                continue
            keep.append(order[i])
            if i + 1 == n:
                break
            w = np.maximum(np.float32(0), np.minimum(x2[i], x2[i + 1:]) - np.maximum(x1[i], x1[i + 1:]))
            h = np.maximum(np.float32(0), np.minimum(y2[i], y2[i + 1:]) - np.maximum(y1[i], y1[i + 1:]))
            inter = w * h
            iou = inter / (area[i] + area[i + 1:] - inter)
            dead[i + 1:] |= iou.astype(np.float64) > float(thr)
    return np.asarray(keep, np.int64)


def greedy_nms(boxes: torch.Tensor, scores: torch.Tensor, thr: float, impl: str = "torchvision") -> torch.Tensor:
    """nms.py:151-156 dispatch.  The reference takes the torchvision branch whenever torchvision is loaded."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    if impl == "torchvision":
        import torchvision

        return torchvision.ops.nms(boxes, scores, thr)
    return torch.from_numpy(greedy_nms_plain(boxes.numpy(), scores.numpy(), thr))


def fast_nms(boxes: torch.Tensor, scores: torch.Tensor, thr: float, iou="probiou") -> torch.Tensor:
    """Fast-NMS, nms.py:187-236 (use_triu branch): keep j iff no higher-ranked row has iou >= thr."""
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64)
    order = stable_desc_order(scores)
    bs = boxes[order]
    m = probiou_matrix(bs, bs) if iou == "probiou" else box_iou_matrix(bs, bs)
    m = m.triu_(diagonal=1)
    ok = torch.nonzero((m >= thr).sum(0) <= 0).squeeze_(-1)
    return order[ok]


def fast_nms_plain(boxes: torch.Tensor, scores: torch.Tensor, thr: float, chunk: int = 2048) -> torch.Tensor:
    """Memory-bounded variant of ``fast_nms`` (same arithmetic per pair, column chunks) for large n."""
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64)
    order = stable_desc_order(scores)
    bs = boxes[order]
    ok = torch.ones(n, dtype=torch.bool)
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        m = probiou_matrix(bs[:c1], bs[c0:c1])
        rows = torch.arange(c1).view(-1, 1)
        cols = torch.arange(c0, c1).view(1, -1)
        ok[c0:c1] = ~((m >= thr) & (rows < cols)).any(0)
    return order[ok]


# ----------------------------------------------------------------------------------------------
# non_max_suppression  (utils/nms.py:13-166)
# ----------------------------------------------------------------------------------------------
def _to_corners(t: torch.Tensor) -> torch.Tensor:
    """ops.py:268-284: xy -/+ wh/2, computed in the tensor's own dtype."""
    half = t[..., 2:] / 2
    return torch.cat((t[..., :2] - half, t[..., :2] + half), -1)


def nms_oracle(pred: torch.Tensor, conf: float = 0.25, iou: float = 0.45, classes=None, agnostic=False,
               multi_label=False, labels=(), max_det=300, nc=0, max_nms=30000, max_wh=7680, rotated=False,
               end2end=False, greedy_impl="torchvision"):
    """Restatement of ``non_max_suppression`` (nms.py:13-166) without the wall-clock guard (nms.py:81,162-164).

    Differences by design, all documented in DESIGN.md: the input is NOT mutated (nms.py:86 rewrites it in
    place); the ``max_nms`` cut and Fast-NMS rank ties by lower row index first (the reference's argsort is
    unstable, nms.py:138,217); the time limit never fires.

    returns (rows, idxs): per image ``(n_i, 6+extra)`` fp32 rows and ``(n_i,)`` int64 anchor indices.
    """
    assert 0 <= conf <= 1 and 0 <= iou <= 1  # nms.py:59-60
    if isinstance(pred, (list, tuple)):
        pred = pred[0]
    cls_keep = None if classes is None else torch.tensor(classes)
    if pred.shape[-1] == 6 or end2end:  # nms.py:66-70
        out = [p[p[:, 4] > conf][:max_det] for p in pred]
        if cls_keep is not None:
            out = [p[(p[:, 5:6] == cls_keep).any(1)] for p in out]
        return out, None
    bs, ch, na = pred.shape
    nc = nc or (ch - 4)
    extra = ch - nc - 4
    cand = pred[:, 4:4 + nc].amax(1) > conf  # nms.py:76 (scalar cast to the tensor dtype)
    multi_label = bool(multi_label) and nc > 1
    rows_all = pred.transpose(-1, -2)
    if not rotated:
        rows_all = torch.cat((_to_corners(rows_all[..., :4]), rows_all[..., 4:]), -1)  # nms.py:86
    aidx = torch.arange(na)
    outs = [torch.zeros((0, 6 + extra))] * bs
    keeps = [torch.zeros((0,), dtype=torch.int64)] * bs
    for b in range(bs):
        x = rows_all[b][cand[b]]
        k = aidx[cand[b]]
        if labels and len(labels[b]) and not rotated:  # nms.py:100-105
            lb = labels[b]
            v = torch.zeros((len(lb), nc + extra + 4))
            v[:, :4] = _to_corners(lb[:, 1:5])
            v[range(len(lb)), lb[:, 0].long() + 4] = 1.0
            x = torch.cat((x, v), 0)
            k = torch.cat((k, torch.full((len(lb),), -1, dtype=torch.int64)))
        if not x.shape[0]:
            continue
        box, cl, ex = x.split((4, nc, extra), 1)
        if multi_label:  # nms.py:114-118
            i, j = torch.where(cl > conf)
            x = torch.cat((box[i], x[i, 4 + j, None], j[:, None].float(), ex[i]), 1)
            k = k[i]
        else:  # nms.py:119-124
            sc, j = cl.max(1, keepdim=True)
            f = sc.view(-1) > conf
            x = torch.cat((box, sc, j.float(), ex), 1)[f]
            k = k[f]
        if cls_keep is not None:  # nms.py:127-131
            f = (x[:, 5:6] == cls_keep).any(1)
            x, k = x[f], k[f]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:  # nms.py:136-141 (stable here)
            f = stable_desc_order(x[:, 4])[:max_nms]
            x, k = x[f], k[f]
        off = x[:, 5:6] * (0 if agnostic else max_wh)  # nms.py:143
        sc = x[:, 4]
        if rotated:
            nb = torch.cat((x[:, :2] + off, x[:, 2:4], x[:, -1:]), dim=-1)  # nms.py:146
            if n > 4096:
                kept = fast_nms_plain(nb, sc, iou)
            else:
                kept = fast_nms(nb, sc, iou, "probiou")
        else:
            kept = greedy_nms(x[:, :4] + off, sc, iou, greedy_impl)  # nms.py:149-156
        kept = kept[:max_det]
        outs[b] = x[kept].float()
        keeps[b] = k[kept]
    return outs, keeps


def postprocess_oracle(levels, strides, nc, conf, iou, reg_max=16, angle_logits=None, **kw):
    """decode + NMS, the composition timed as the CPU baseline."""
    if angle_logits is None:
        y = decode_oracle(levels, strides, nc, reg_max)
        return nms_oracle(y, conf, iou, nc=nc, **kw)
    y = obb_forward_oracle(levels, angle_logits, strides, nc, reg_max)
    return nms_oracle(y, conf, iou, nc=nc, rotated=True, **kw)
