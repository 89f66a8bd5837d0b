"""Drop-ins for the validator step that follows NMS:

  match_predictions(self, pred_classes, true_classes, iou, use_scipy=False)   engine/validator.py:267-307
  process_batch(self, preds, batch)                                           models/yolo/detect/val.py:274-288
  match_batch(...)                                                            the same for every image of a batch, one launch

The reference moves the IoU matrix to the host and runs numpy nonzero / argsort / unique per IoU level; here one CTA per
image computes the (N, niou) true-positive matrix on the device (``match_predictions_kernel``), box_iou (metrics.py:54)
included.  ``use_scipy=True`` (Hungarian matching) is not built - ``patch.install`` leaves it with the reference.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi, engine


_thr_cache: dict = {}


def _thresholds(iouv):
    """Host copy of the validator's IoU levels (``self.iouv``, a device tensor: validator.py:148) - read back once per
    tensor, not on every call."""
    key = None
    if isinstance(iouv, torch.Tensor):
        key = (id(iouv), iouv.data_ptr(), iouv.numel())
        hit = _thr_cache.get(key)
        if hit is not None and hit[0] is iouv:
            return hit[1], hit[2]
    vals = [float(v) for v in (iouv.cpu().tolist() if isinstance(iouv, torch.Tensor) else iouv)]
    if not 1 <= len(vals) <= 16:
        raise ValueError(f"{len(vals)} IoU levels: the kernel takes 1..16")
    arr = (C.c_float * len(vals))(*vals)
    if key is not None:
        if len(_thr_cache) > 64:
            _thr_cache.clear()
        _thr_cache[key] = (iouv, arr, len(vals))
    return arr, len(vals)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t if t.dtype == torch.float32 else t.float()
    return t if t.is_contiguous() else t.contiguous()


def match_iou_matrix(iouv, pred_classes: torch.Tensor, true_classes: torch.Tensor, iou: torch.Tensor) -> torch.Tensor:
    """(N, niou) bool true-positive matrix from an (M, N) IoU matrix (labels x detections), validator.py:267-307."""
    _cabi.require_cuda(iou, "match_predictions")
    n, m = pred_classes.shape[0], true_classes.shape[0]
    thr, nthr = _thresholds(iouv)
    correct = torch.zeros((n, nthr), dtype=torch.uint8, device=iou.device)
    if n and m:
        pc, tc, mat = _f32c(pred_classes.to(iou.device)), _f32c(true_classes.to(iou.device)), _f32c(iou)
        need = nthr * m * 4
        ws = engine._scratch(iou.device, need) if need > 200 * 1024 else None
        rc = _cabi.load().ypb_match_predictions(pc.data_ptr(), 0, 1, 0, 1, n, None, None, None, m, m, mat.data_ptr(),
                                                mat.stride(0), tc.data_ptr(), thr, nthr, correct.data_ptr(),
                                                ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0,
                                                _cabi.stream_ptr(iou.device))
        _cabi.check(rc, "ypb_match_predictions")
    return correct.bool()


def match_predictions(self, pred_classes, true_classes, iou, use_scipy: bool = False):
    """``BaseValidator.match_predictions`` drop-in (validator.py:267-307); reads ``self.iouv``."""
    if use_scipy:
        raise NotImplementedError("use_scipy=True (linear_sum_assignment) is not built into the kernel")
    return match_iou_matrix(self.iouv, pred_classes, true_classes, iou)


def match_boxes(iouv, pred_boxes: torch.Tensor, pred_cls: torch.Tensor, gt_boxes: torch.Tensor, gt_cls: torch.Tensor):
    """box_iou (metrics.py:54) + match_predictions for one image without materialising the IoU matrix."""
    _cabi.require_cuda(pred_boxes, "match_boxes")
    dev = pred_boxes.device
    n, m = pred_boxes.shape[0], gt_boxes.shape[0]
    thr, nthr = _thresholds(iouv)
    correct = torch.zeros((n, nthr), dtype=torch.uint8, device=dev)
    if n and m:
        rows = torch.cat([_f32c(pred_boxes), _f32c(pred_cls.to(dev)).view(-1, 1)], 1)
        labels = torch.cat([_f32c(gt_cls.to(dev)).view(-1, 1), _f32c(gt_boxes.to(dev))], 1)
        need = nthr * m * 4
        ws = engine._scratch(dev, need) if need > 200 * 1024 else None
        rc = _cabi.load().ypb_match_predictions(rows.data_ptr(), 0, 5, 4, 1, n, None, labels.data_ptr(), None, m, m, None, 0,
                                                None, thr, nthr, correct.data_ptr(),
                                                ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0,
                                                _cabi.stream_ptr(dev))
        _cabi.check(rc, "ypb_match_predictions")
    return correct.bool()


def process_batch(self, preds: dict, batch: dict) -> dict:
    """``DetectionValidator._process_batch`` drop-in (detect/val.py:274-288): {"tp": (N, niou) bool ndarray}."""
    import numpy as np

    if batch["cls"].shape[0] == 0 or preds["cls"].shape[0] == 0:
        return {"tp": np.zeros((preds["cls"].shape[0], self.niou), dtype=bool)}
    return {"tp": match_boxes(self.iouv, preds["bboxes"], preds["cls"], batch["bboxes"], batch["cls"]).cpu().numpy()}


def match_batch(iouv, rows: torch.Tensor, count: torch.Tensor | None, labels: torch.Tensor, label_counts) -> torch.Tensor:
    """True-positive matrices of a whole validation batch in one launch.

    rows (B, max_det, >=6) NMS result rows (x1,y1,x2,y2,conf,cls) and their device ``count``; labels (sum M, 5) fp32
    cls,x1,y1,x2,y2 of all images in the rows' coordinate frame, ``label_counts`` per-image label counts (host).
    Returns (B, max_det, niou) uint8 (rows past ``count`` are zero)."""
    _cabi.require_cuda(rows, "match_batch")
    if rows.dtype != torch.float32 or rows.dim() != 3 or rows.shape[2] < 6 or rows.stride(2) != 1:
        raise ValueError("rows must be (B, max_det, >=6) float32 with contiguous columns")
    b, md, _ = rows.shape
    if len(label_counts) != b:
        raise ValueError("label_counts does not match the batch")
    thr, nthr = _thresholds(iouv)
    offs = [0]
    for c in label_counts:
        offs.append(offs[-1] + int(c))
    correct = torch.zeros((b, md, nthr), dtype=torch.uint8, device=rows.device)
    if b == 0 or md == 0 or offs[-1] == 0:
        return correct
    labels = _f32c(labels.to(rows.device))
    if labels.shape != (offs[-1], 5):
        raise ValueError(f"labels must be ({offs[-1]}, 5), got {tuple(labels.shape)}")
    offsets = torch.tensor(offs, dtype=torch.int32, device=rows.device)
    max_m = max(int(c) for c in label_counts)
    need = nthr * offs[-1] * 4
    ws = engine._scratch(rows.device, need) if nthr * max_m * 4 > 200 * 1024 else None
    rc = _cabi.load().ypb_match_predictions(rows.data_ptr(), rows.stride(0), rows.stride(1), 5, b, md,
                                            count.data_ptr() if count is not None else None, labels.data_ptr(),
                                            offsets.data_ptr(), 0, max_m, None, 0, None, thr, nthr, correct.data_ptr(),
                                            ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0,
                                            _cabi.stream_ptr(rows.device))
    _cabi.check(rc, "ypb_match_predictions")
    return correct
