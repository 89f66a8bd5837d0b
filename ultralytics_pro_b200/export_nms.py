"""The exporter's embedded NMS (``engine/exporter.py:1389-1481`` ``NMSModel.forward``, everything after ``self.model(x)``)
as one sync-free device call.

Differences from ``non_max_suppression`` that this mirrors exactly: boxes arrive as corners (the export decode,
head.py:189 with ``xyxy=True``); one row per anchor (``scores.max``); suppression runs on boxes normalised by the larger
image side and scaled by ``1/nc``, with the class offset in the same unit (exporter.py:1437-1452 - the rounding of every IoU
differs from the ``max_wh`` flavour); the result is a fixed-size ``(B, max_det, 6+extra)`` tensor, zero-padded, no host sync.
Axis-aligned tasks (detect / segment / pose); the OBB branch (``fast_nms`` with ``exit_early=False``) stays with the reference.
"""
from __future__ import annotations

import torch

from . import _cabi, engine
from .nms import _greedy_threshold


def nms_model_postprocess(pred: torch.Tensor, image_hw, nc: int, conf: float, iou: float, max_det: int,
                          agnostic_nms: bool = False, return_count: bool = False):
    """pred: (B, 4+nc+extra, A) float CUDA tensor, boxes xyxy in pixels of the ``image_hw`` network input.
    Returns (B, min(max_det, A), 6+extra) fp32: x1,y1,x2,y2,score,cls,extras per kept detection in descending-score order,
    zeros after the last one (and, with ``return_count``, the device int32 counts)."""
    _cabi.require_cuda(pred, "nms_model_postprocess")
    if pred.dim() != 3:
        raise ValueError(f"pred must be (B, 4+nc+extra, A), got {tuple(pred.shape)}")
    b, ch, a = pred.shape
    extra = ch - 4 - nc
    if nc < 1 or extra < 0:
        raise ValueError(f"nc={nc} inconsistent with {ch} channels")
    max_det = min(a, int(max_det))  # exporter.py:1431
    mult = _cabi.f32_round(1.0 / max(nc, 1))  # exporter.py:1441 (non-OBB)
    div = _cabi.f32_round(float(max(image_hw)))  # exporter.py:1444 torch.tensor(x.shape[2:]).max()
    plan = engine.make_plan(pred.device, b, a, nc, extra, _cabi.round_to_dtype(float(conf), pred.dtype),
                            _greedy_threshold(iou), max_det, a, 0.0 if agnostic_nms else mult, False, _cabi.RULE_GREEDY,
                            None, nms_box=(div, mult), boxes_xyxy=True, pad_output=True)
    if b:
        engine.run_from_dense(pred, plan)
    else:
        plan.count.zero_()
    return (plan.rows, plan.count) if return_count else plan.rows
