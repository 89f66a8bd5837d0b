"""Drop-in for ``ultralytics/utils/nms.py``: ``non_max_suppression`` and ``TorchNMS`` with the reference's exact
signatures, executed by the sm_100a kernels of libyolopost_b200 (no CPU path, no torch arithmetic).

Behavioural notes against the reference (all stated in DESIGN.md):
  * the input tensor is not modified (the reference rewrites its box channels in place, nms.py:86) unless
    ``MUTATE_INPUT_LIKE_REFERENCE`` is set;
  * ``max_time_img`` is accepted and ignored - the wall-clock guard (nms.py:81,162-164) silently drops images;
  * score ties rank by lower row first everywhere (== torchvision's stable sort; the reference's own ``argsort`` at
    nms.py:138,217,264 is unstable, so its tie order is undefined);
  * an image with no detections yields a ``(0,)`` index tensor with ``return_idxs`` (reference: ``(0, 1)``).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi, engine

# "torchvision": suppress iff fp32 IoU, widened to double, > iou_thres (torchvision's CPU kernel - the branch the
# reference takes whenever torchvision is imported, nms.py:151-154).  "torchnms": fp32 compare against float32(iou_thres)
# (TorchNMS.nms, nms.py:294).  They differ only when float32(iou_thres) > iou_thres and an IoU equals it exactly.
GREEDY_THRESHOLD_SEMANTICS = "torchvision"
MUTATE_INPUT_LIKE_REFERENCE = False


def _greedy_threshold(iou_thres: float) -> float:
    if GREEDY_THRESHOLD_SEMANTICS == "torchvision":
        return _cabi.largest_f32_not_above(float(iou_thres))
    return _cabi.f32_round(float(iou_thres))


def _end2end_select(prediction, conf_thres, max_det, classes):
    """nms.py:66-70: (B, N, 6) end-to-end output - threshold, cap, optional class filter.  Index plumbing only."""
    pred = prediction
    keep = pred[..., 4] > conf_thres
    rank = keep.cumsum(1)
    keep &= rank <= max_det
    if classes is not None:
        cls = torch.as_tensor(classes, device=pred.device)
        keep &= (pred[..., 5:6] == cls).any(-1)
    order = torch.sort((~keep).to(torch.int8), dim=1, stable=True).indices
    gathered = pred.gather(1, order.unsqueeze(-1).expand(-1, -1, pred.shape[-1]))
    counts = engine.fetch_counts(keep.sum(1).to(torch.int32))
    return [gathered[b, :n] for b, n in enumerate(counts)]


def _append_labels(prediction, labels, nc, extra):
    """nms.py:100-105 (autolabelling): a-priori boxes become extra anchors with a one-hot score of 1.0."""
    b, ch, _ = prediction.shape
    lmax = max(len(l) for l in labels)
    add = torch.zeros((b, ch, lmax), dtype=prediction.dtype, device=prediction.device)
    for i, lb in enumerate(labels):
        if len(lb):
            lb = torch.as_tensor(lb, device=prediction.device).to(torch.float32)
            n = lb.shape[0]
            add[i, :4, :n] = lb[:, 1:5].t().to(prediction.dtype)
            add[i, lb[:, 0].long() + 4, torch.arange(n, device=prediction.device)] = 1.0
    return torch.cat((prediction, add), 2)


def non_max_suppression(
    prediction,
    conf_thres: float = 0.25,
    iou_thres: float = 0.45,
    classes=None,
    agnostic: bool = False,
    multi_label: bool = False,
    labels=(),
    max_det: int = 300,
    nc: int = 0,  # number of classes (optional)
    max_time_img: float = 0.05,
    max_nms: int = 30000,
    max_wh: int = 7680,
    rotated: bool = False,
    end2end: bool = False,
    return_idxs: bool = False,
):
    """Non-maximum suppression on a decoded prediction tensor; same contract as the reference (nms.py:13-166).

    Args mirror the reference one for one.  ``prediction`` is (B, 4+nc+extra, A) float32/float16/bfloat16 on a CUDA
    device (any strides).  Returns ``list[Tensor(n_i, 6+extra)]`` fp32 rows ``x1,y1,x2,y2,conf,cls,extra...``
    (``cx,cy,w,h,conf,cls,angle`` when ``rotated``) in descending-score order, plus the kept anchor indices when
    ``return_idxs``.
    """
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if isinstance(prediction, (list, tuple)):  # (inference_out, loss_out), nms.py:61
        prediction = prediction[0]
    _cabi.require_cuda(prediction, "non_max_suppression")
    from . import lazy

    if isinstance(prediction, lazy.LazyDecoded):
        # the tensor came from our Detect._inference drop-in and nobody has needed its values yet: run the fused
        # head -> NMS kernels on the recorded level tensors (bit-identical to decode + NMS); the dense tensor is never written
        rec = prediction.head_record()
        if (rec is not None and not rotated and not end2end and not rec.xyxy and nc in (0, rec.nc)
                and not (labels and any(len(l) for l in labels)) and not MUTATE_INPUT_LIKE_REFERENCE):
            from .head import postprocess_from_head

            lazy.STATS["fused"] += 1
            return postprocess_from_head(rec.levels, rec.strides, rec.nc, conf_thres, iou_thres, classes, agnostic,
                                         multi_label, max_det, max_nms, max_wh, rec.reg_max, return_idxs=return_idxs)
        prediction = prediction.materialize()
    if prediction.shape[-1] == 6 or end2end:
        return _end2end_select(prediction, conf_thres, max_det, classes)

    if prediction.dim() != 3:
        raise ValueError(f"prediction must be (B, 4+nc+extra, A), got {tuple(prediction.shape)}")
    bs, ch, _ = prediction.shape
    nc = nc or (ch - 4)
    extra = ch - nc - 4
    if nc < 1 or extra < 0:
        raise ValueError(f"nc={nc} inconsistent with {ch} channels")
    multi_label = bool(multi_label) and nc > 1
    if labels and any(len(l) for l in labels) and not rotated:
        prediction = _append_labels(prediction, labels, nc, extra)
    na = prediction.shape[2]
    if bs == 0 or na == 0:
        empty = [torch.zeros((0, 6 + extra), device=prediction.device)] * bs
        return (empty, [torch.zeros((0,), dtype=torch.int64, device=prediction.device)] * bs) if return_idxs else empty

    dt = prediction.dtype
    conf_t = _cabi.round_to_dtype(float(conf_thres), dt)
    if rotated:
        rule, iou_eff = _cabi.RULE_FAST_PROBIOU, _cabi.f32_round(float(iou_thres))
    else:
        rule, iou_eff = _cabi.RULE_GREEDY, _greedy_threshold(iou_thres)
    plan = engine.make_plan(prediction.device, bs, na, nc, extra, conf_t, iou_eff, max_det, max_nms,
                            0.0 if agnostic else float(max_wh), multi_label, rule, classes, cached=True)
    engine.run_from_dense(prediction, plan)
    if MUTATE_INPUT_LIKE_REFERENCE and not rotated:
        xy, wh = prediction[:, :2].clone(), prediction[:, 2:4] / 2
        prediction[:, :2], prediction[:, 2:4] = xy - wh, xy + wh
    return engine.split_results(plan, return_idxs)


def _pairwise(box1: torch.Tensor, box2: torch.Tensor, dim: int, what: str) -> torch.Tensor:
    _cabi.require_cuda(box1, what)
    if box1.dim() != 2 or box2.dim() != 2 or box1.shape[1] != dim or box2.shape[1] != dim:
        raise ValueError(f"{what}: expected (N, {dim}) and (M, {dim}), got {tuple(box1.shape)} and {tuple(box2.shape)}")
    a = box1.to(torch.float32).contiguous()
    b = box2.to(device=box1.device, dtype=torch.float32).contiguous()
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=box1.device)
    rc = _cabi.load().ypb_pairwise_iou(a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], dim, out.data_ptr(),
                                       _cabi.stream_ptr(box1.device))
    _cabi.check(rc, "ypb_pairwise_iou")
    return out


def box_iou(box1, box2, eps: float = 1e-7):
    """``metrics.box_iou`` (metrics.py:54-75): (N, 4), (M, 4) xyxy -> (N, M) fp32.  As ``iou_func`` of ``TorchNMS.fast_nms`` it is
    recognised by name and evaluated pair by pair inside the suppression kernel (the matrix is never materialised)."""
    if eps != 1e-7:
        raise NotImplementedError("box_iou: the kernel is built with the reference's default eps=1e-7")
    return _pairwise(box1, box2, 4, "box_iou")


def batch_probiou(obb1, obb2, eps: float = 1e-7):
    """``metrics.batch_probiou`` (metrics.py:251-284): (N, 5), (M, 5) xywhr -> (N, M) fp32; same remark as ``box_iou``."""
    if eps != 1e-7:
        raise NotImplementedError("batch_probiou: the kernel is built with the reference's default eps=1e-7")
    return _pairwise(obb1, obb2, 5, "batch_probiou")


def _nms_boxes(boxes: torch.Tensor, scores: torch.Tensor, rule: int, thr: float) -> torch.Tensor:
    _cabi.require_cuda(boxes, "TorchNMS")
    n = boxes.shape[0]
    boxes = boxes.to(torch.float32).contiguous()
    scores = scores.to(torch.float32).contiguous()
    lib = _cabi.load()
    nbytes = lib.ypb_nms_boxes_workspace_bytes(n)
    scratch = engine._scratch(boxes.device, nbytes)
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=boxes.device)
    count = torch.empty((1,), dtype=torch.int32, device=boxes.device)
    rc = lib.ypb_nms_boxes(boxes.data_ptr(), scores.data_ptr(), n, boxes.shape[1], rule, thr, keep.data_ptr(),
                           count.data_ptr(), scratch.data_ptr(), scratch.numel(), _cabi.stream_ptr(boxes.device))
    _cabi.check(rc, "ypb_nms_boxes")
    return keep[: engine.fetch_counts(count)[0]]


class TorchNMS:
    """Mirror of the reference's ``TorchNMS`` (nms.py:169-337)."""

    @staticmethod
    def fast_nms(boxes, scores, iou_threshold: float, use_triu: bool = True, iou_func=box_iou, exit_early: bool = True):
        """Fast-NMS (nms.py:187-236): keep a box iff no higher-scoring box, kept or not, overlaps it >= threshold."""
        if boxes.numel() == 0 and exit_early:
            return torch.empty((0,), dtype=torch.int64, device=boxes.device)
        if not use_triu:
            raise NotImplementedError("use_triu=False is the reference's export-graph branch (nms.py:224-235)")
        name = getattr(iou_func, "__name__", "")
        if name == "batch_probiou":
            rule = _cabi.RULE_FAST_PROBIOU
        elif name == "box_iou":
            rule = _cabi.RULE_FAST_BOXIOU
        else:
            raise NotImplementedError(f"iou_func {name!r}: only box_iou and batch_probiou are built into the kernel")
        return _nms_boxes(boxes, scores, rule, _cabi.f32_round(float(iou_threshold)))

    @staticmethod
    def nms(boxes, scores, iou_threshold: float):
        """Greedy NMS (nms.py:239-296); indices of kept boxes in descending-score order."""
        if boxes.numel() == 0:
            return torch.empty((0,), dtype=torch.int64, device=boxes.device)
        return _nms_boxes(boxes, scores, _cabi.RULE_GREEDY, _cabi.f32_round(float(iou_threshold)))

    @staticmethod
    def batched_nms(boxes, scores, idxs, iou_threshold: float, use_fast_nms: bool = False):
        """Class-aware NMS by coordinate offset (nms.py:299-337)."""
        if boxes.numel() == 0:
            return torch.empty((0,), dtype=torch.int64, device=boxes.device)
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + 1)
        shifted = boxes + offsets[:, None]
        if use_fast_nms:
            return TorchNMS.fast_nms(shifted, scores, iou_threshold)
        return TorchNMS.nms(shifted, scores, iou_threshold)
