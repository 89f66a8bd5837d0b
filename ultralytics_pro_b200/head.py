"""Drop-ins for the Detect-family decode (``ultralytics/nn/modules/head.py``) and the fused post-process call.

  detect_inference(self, x)      bound as ``Detect._inference`` (head.py:151-169; identical copies :340,:535,:724)
  detect_decode_bboxes / obb_decode_bboxes / dfl_forward
                                 bound as ``Detect.decode_bboxes`` (head.py:184-191), ``OBB.decode_bboxes`` (:1040-1042) and
                                 ``DFL.forward`` (block.py:250-253) - the pieces ``YOLOEDetect.forward_lrpc`` (:1777-1813) calls
  decode_head(...)               the same decode as a free function on raw level tensors
  postprocess_from_head(...)     decode + non_max_suppression in one pass over the head (never writes the dense
                                 tensor) - the `postprocess` envelope of predictor.py:335 / validator.py:221

All arithmetic runs in libyolopost_b200; PyTorch allocates the output and provides the stream.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi, engine
from .nms import _greedy_threshold


def _anchor_cache(levels, strides, dtype, device):
    """anchors (2, A) / strides (1, A) exactly as head.py:163-165 caches them (Pose.kpts_decode and
    BaseModel._apply read these attributes).  Built once per input shape with index arithmetic only."""
    pts, srow = [], []
    for lv, s in zip(levels, strides):
        h, w = lv.shape[2], lv.shape[3]
        gx = torch.arange(w, device=device, dtype=dtype) + 0.5
        gy = torch.arange(h, device=device, dtype=dtype) + 0.5
        yy, xx = torch.meshgrid(gy, gx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        srow.append(torch.full((h * w, 1), float(s), dtype=dtype, device=device))
    return torch.cat(pts).transpose(0, 1), torch.cat(srow).transpose(0, 1)


def decode_head(levels, strides, nc: int, reg_max: int = 16, angle: torch.Tensor | None = None,
                angle_is_logit: bool = False, append_angle: bool = False, xyxy: bool = False) -> torch.Tensor:
    """Dense decode of raw Detect-head levels -> (B, 4+nc[+1], A) in the input dtype.

    levels: list of (B, 4*reg_max+nc, H_i, W_i) CUDA tensors; strides: per-level model stride (head.py:79);
    angle: optional (B, 1, A) OBB channel - raw cv4 logits when ``angle_is_logit`` else already
    ``(sigmoid-0.25)*pi`` (head.py:1031) - selecting the rotated decode (tal.py:385-403).
    """
    desc, keep, anchors = engine.head_desc(levels, strides, nc, reg_max)
    lv0 = keep[0]
    b = lv0.shape[0]
    ch = 4 + nc + (1 if (angle is not None and append_angle) else 0)
    out = torch.empty((b, ch, anchors), dtype=lv0.dtype, device=lv0.device)
    ang_ptr = None
    if angle is not None:
        _cabi.require_cuda(angle, "angle")
        if angle.dtype != lv0.dtype:
            angle = angle.to(lv0.dtype)
        angle = angle.reshape(b, anchors).contiguous()
        ang_ptr = angle.data_ptr()
    lib = _cabi.load()
    rc = lib.ypb_decode_dense(C.byref(desc), ang_ptr, int(angle_is_logit), int(append_angle), int(xyxy),
                              out.data_ptr(), desc.dtype, out.stride(0), out.stride(1), _cabi.stream_ptr(lv0.device))
    _cabi.check(rc, "ypb_decode_dense")
    return out


def host_strides(self) -> tuple:
    """``self.stride`` (head.py:79; a device tensor once ``BaseModel._apply`` moved the model) as host floats, read back
    ONCE per stride tensor: ``float(s) for s in self.stride`` would be one device->host sync per level on every forward -
    the reference's ``_inference`` has none on a cache hit (``make_anchors`` reads the strides only when the shape changes)."""
    st = self.stride
    key = (id(st), st.data_ptr() if isinstance(st, torch.Tensor) else None, getattr(st, "device", None))
    cached = self.__dict__.get("_ypb_stride_cache")
    if cached is None or cached[0] != key:
        vals = tuple(float(v) for v in (st.tolist() if isinstance(st, torch.Tensor) else st))
        cached = (key, vals, st)  # holding `st` keeps id() unique
        self.__dict__["_ypb_stride_cache"] = cached
    return cached[1]


def _is_obb(self) -> bool:
    return hasattr(self, "ne") and hasattr(self, "cv4") and getattr(self, "angle", None) is not None


def _has_riders(self) -> bool:
    """Segment / Pose / OBB heads concatenate their own channels to ``_inference``'s result right away (head.py:837,
    :1038, :1252), so a lazily materialised tensor would only add overhead there."""
    return any(hasattr(self, a) for a in ("nm", "ne", "kpt_shape", "nk"))


def detect_inference(self, x: list) -> torch.Tensor:
    """``Detect._inference`` drop-in (head.py:151-169).  Reads the same module attributes, keeps caching
    ``self.anchors/self.strides/self.shape`` and returns the dense (B, 4+nc, A) tensor in the input dtype.

    With ``lazy.ENABLED`` (set by ``patch.install(lazy_decode=True)``) a plain Detect head returns a ``LazyDecoded``
    tensor instead: same shape/dtype/device, materialised by the dense kernel on first use by ANY torch op - unless the
    first consumer is the patched ``non_max_suppression``, which then runs the fused head->NMS kernels on the level
    tensors and the dense tensor is never written (SURVEY.md section 7 "Drop-in surface")."""
    shape = x[0].shape  # BCHW
    self._ypb_level_hw = [(int(lv.shape[2]), int(lv.shape[3])) for lv in x]  # read by pose_kpts_decode
    strides = host_strides(self)
    if self.dynamic or self.shape != shape:
        self.anchors, self.strides = _anchor_cache(x, strides, x[0].dtype, x[0].device)
        self.shape = shape
    angle = self.angle if _is_obb(self) else None  # OBB family (head.py:1032)
    xyxy = bool(self.end2end or self.xyxy)
    from . import lazy

    if lazy.ENABLED and angle is None and not self.end2end and not _has_riders(self) and not torch.is_grad_enabled():
        return lazy.LazyDecoded.from_head(list(x), strides, self.nc, self.reg_max, xyxy)
    return decode_head(x, strides, self.nc, self.reg_max, angle=angle, angle_is_logit=False, append_angle=False, xyxy=xyxy)


def dfl_forward(self, x: torch.Tensor) -> torch.Tensor:
    """``DFL.forward`` drop-in (block.py:250-253): (B, 4*c1, A) -> (B, 4, A), softmax expectation over the c1 bins of
    every side (the frozen ``arange(c1)`` 1x1 conv of block.py:245-247 folded into the kernel)."""
    _cabi.require_cuda(x, "DFL.forward")
    b, ch, a = x.shape
    if x.stride(2) != 1 and a > 1:
        x = x.contiguous()
    out = torch.empty((b, 4, a), dtype=x.dtype, device=x.device)
    rc = _cabi.load().ypb_dfl_expectation(x.data_ptr(), _cabi.dtype_code(x.dtype), b, ch // 4, a, x.stride(0), x.stride(1),
                                          out.data_ptr(), out.stride(0), out.stride(1), _cabi.stream_ptr(x.device))
    _cabi.check(rc, "ypb_dfl_expectation")
    return out


def dist2bbox(distance: torch.Tensor, anchor_points: torch.Tensor, xywh: bool = True, angle: torch.Tensor | None = None):
    """``dist2bbox(distance, anchor_points, xywh, dim=1)`` (tal.py:367-376) or, with ``angle``, ``dist2rbox(distance,
    angle, anchor_points, dim=1)`` (tal.py:385-403): distance (B, 4, A), anchor_points (2, A) / (1|B, 2, A) of any
    strides, angle (B, 1, A) -> (B, 4, A) in the input dtype."""
    _cabi.require_cuda(distance, "decode_bboxes")
    if distance.dim() != 3 or distance.shape[1] != 4:
        raise ValueError(f"distance must be (B, 4, A), got {tuple(distance.shape)}")
    b, _, a = distance.shape
    dt = distance.dtype
    if distance.stride(2) != 1 and a > 1:
        distance = distance.contiguous()
    ap = anchor_points if anchor_points.dim() == 3 else anchor_points.unsqueeze(0)
    if ap.shape[-2:] != (2, a) or ap.shape[0] not in (1, b):
        raise ValueError(f"anchor_points {tuple(anchor_points.shape)} does not broadcast against (B={b}, 2, A={a})")
    if ap.dtype != dt or ap.device != distance.device:
        ap = ap.to(device=distance.device, dtype=dt)
    ang_ptr, ang_sb = None, 0
    if angle is not None:
        angle = angle.to(dt).reshape(b, a)
        if not angle.is_contiguous():
            angle = angle.contiguous()
        ang_ptr, ang_sb = angle.data_ptr(), angle.stride(0)
    out = torch.empty((b, 4, a), dtype=dt, device=distance.device)
    rc = _cabi.load().ypb_dist2bbox(distance.data_ptr(), distance.stride(0), distance.stride(1), ap.data_ptr(),
                                    ap.stride(0) if ap.shape[0] == b and b > 1 else 0, ap.stride(1), ap.stride(2), ang_ptr,
                                    ang_sb, _cabi.dtype_code(dt), b, a, int(bool(xywh)), out.data_ptr(), out.stride(0),
                                    out.stride(1), _cabi.stream_ptr(distance.device))
    _cabi.check(rc, "ypb_dist2bbox")
    return out


def detect_decode_bboxes(self, bboxes: torch.Tensor, anchors: torch.Tensor, xywh: bool = True) -> torch.Tensor:
    """``Detect.decode_bboxes`` drop-in (head.py:184-191)."""
    return dist2bbox(bboxes, anchors, xywh=xywh and not self.end2end and not self.xyxy)


def obb_decode_bboxes(self, bboxes: torch.Tensor, anchors: torch.Tensor) -> torch.Tensor:
    """``OBB.decode_bboxes`` drop-in (head.py:1040-1042): rotated decode with the angle ``OBB.forward`` stored."""
    return dist2bbox(bboxes, anchors, angle=self.angle)


def decode_keypoints(kpts: torch.Tensor, level_hw, strides, kpt_shape) -> torch.Tensor:
    """``Pose.kpts_decode`` (head.py:1254-1273, non-export branch) on raw (B, nk*ndim, A) keypoint logits:
    x,y -> (v*2 + (anchor - 0.5)) * stride, visibility -> sigmoid.  Returns a new tensor of the same shape and dtype."""
    _cabi.require_cuda(kpts, "decode_keypoints")
    b, ch, a = kpts.shape
    nk, ndim = int(kpt_shape[0]), int(kpt_shape[1])
    if ch != nk * ndim or a != sum(int(h) * int(w) for h, w in level_hw):
        raise ValueError(f"kpts {tuple(kpts.shape)} inconsistent with kpt_shape {tuple(kpt_shape)} / levels {list(level_hw)}")
    if kpts.stride(2) != 1:
        kpts = kpts.contiguous()
    out = torch.empty((b, ch, a), dtype=kpts.dtype, device=kpts.device)
    desc = engine.geometry_desc(level_hw, strides, b, kpts.dtype)
    rc = _cabi.load().ypb_kpts_decode(C.byref(desc), kpts.data_ptr(), kpts.stride(0), kpts.stride(1), ch, ndim,
                                      out.data_ptr(), _cabi.stream_ptr(kpts.device))
    _cabi.check(rc, "ypb_kpts_decode")
    return out


def pose_kpts_decode(self, bs: int, kpts: torch.Tensor) -> torch.Tensor:
    """``Pose.kpts_decode`` drop-in (head.py:1254; identical copies :1322, :1390, :1459).  The level grid sizes come from
    the level list seen by ``detect_inference`` (``Pose.forward`` runs ``Detect.forward`` first, head.py:1249-1252)."""
    hw = getattr(self, "_ypb_level_hw", None)
    strides = host_strides(self)
    if hw is None or sum(h * w for h, w in hw) != kpts.shape[-1]:
        # _inference ran through the reference (not patched): recover the grids from the cached anchor rows
        # (head.py:163-165), once per cached shape - this reads back from the device
        cached = self.__dict__.get("_ypb_hw_from_anchors")
        if cached is None or cached[0] != (tuple(self.shape), kpts.shape[-1]):
            srow = self.strides.view(-1)
            hw = []
            for s in strides:
                sel = srow == s
                w = int(self.anchors[0, sel].max().item() + 0.5)
                hw.append((int(sel.sum()) // w, w))
            cached = ((tuple(self.shape), kpts.shape[-1]), hw)
            self.__dict__["_ypb_hw_from_anchors"] = cached
        hw = cached[1]
    return decode_keypoints(kpts.view(bs, self.nk, -1), hw, strides, self.kpt_shape)


def detect_postprocess(preds: torch.Tensor, max_det: int, nc: int = 80) -> torch.Tensor:
    """``Detect.postprocess`` drop-in (head.py:193-214, the end2end / v10 top-k): preds (B, A, 4+nc) -> (B, K, 6) rows
    ``box(4 as given), score, class`` of the K = min(max_det, A) best (anchor, class) pairs, best first.

    The reference takes the K anchors with the largest class maximum and then the K best pairs among their K*nc scores;
    for tie-free scores that is the global top-K over all pairs (an anchor outside the first set cannot own a pair that
    beats K anchors' maxima).  ONE pass over the scores: the filter + sort kernels rank the anchors by their maximum and yield
    the K best anchors and the K-th value T; a second, tiny call ranks the pairs with score >= T of those K anchors only.
    Index arithmetic between the two (picking T, nextafter, sorting / gathering B x K anchor rows) is torch plumbing on the
    device; nothing is read back."""
    _cabi.require_cuda(preds, "detect_postprocess")
    if preds.dim() != 3 or preds.shape[2] != 4 + nc:
        raise ValueError(f"preds must be (B, A, {4 + nc}), got {tuple(preds.shape)}")
    b, a, _ = preds.shape
    k = min(int(max_det), a)
    dev = preds.device
    if b == 0 or k == 0:
        return torch.zeros((b, k, 6), dtype=preds.dtype, device=dev)
    dense = preds.transpose(1, 2)  # (B, 4+nc, A) view; the kernels take any strides
    ninf = float("-inf")
    # pass 1 (the only pass over all the scores): rank the anchors by their class maximum, keep the K best (head.py:208)
    first = engine.make_plan(dev, b, a, nc, 0, ninf, 1.0, k, a, 0.0, False, _cabi.RULE_GREEDY, boxes_xyxy=True,
                             pad_output=True)  # fewer than K rankable anchors (NaN scores): zero rows, index -1
    engine.run_from_dense(dense, first)
    kth = first.rows[:, k - 1, 4].contiguous()
    thr = torch.nextafter(kth, torch.full_like(kth, ninf))  # score > thr  <=>  score >= K-th anchor maximum
    # pass 2 reads ONLY those K anchors (head.py:209-211 gathers them too): every pair that can reach the result lives in an
    # anchor whose maximum is >= T, and among anchors tied at T the lower indices - the ones pass 1 kept - win the flat-index
    # tie-break, so restricting to the kept anchors is exact.  The kernel reads them where they lie, through the index list
    # pass 1 left on the device (``anchor_subset``): row ids stay those of the full tensor, so "lower row first" is "lower flat
    # index first" with no sorting or gathering in between.  A threshold of 1 makes both suppression calls pure rankings.
    sel = first.idx if first.idx.shape[1] == k else first.idx[:, :k].contiguous()
    second = engine.make_plan(dev, b, a, nc, 0, 0.0, 1.0, k, k * nc, 0.0, True, _cabi.RULE_GREEDY, boxes_xyxy=True,
                              conf_per_image=thr, rows_cap=k * nc)
    engine.run_from_dense(dense, second, anchor_subset=sel)
    return second.rows.to(preds.dtype)


def postprocess_from_head(levels, strides, nc: int, conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None,
                          agnostic: bool = False, multi_label: bool = False, max_det: int = 300, max_nms: int = 30000,
                          max_wh: int = 7680, reg_max: int = 16, angle_logits: torch.Tensor | None = None,
                          return_idxs: bool = False, sync: bool = True, img_shape=None, orig_shapes=None,
                          ratio_pads=None, mask_coeffs: torch.Tensor | None = None,
                          kpt_logits: torch.Tensor | None = None, kpt_shape=None):
    """Fused ``Detect._inference`` + ``non_max_suppression`` (head.py:151-169 then nms.py:13-166).

    Bit-identical to ``non_max_suppression(decode_head(levels, ...), ...)`` but reads the head once and never writes
    the (B, 4+nc, A) tensor.  ``angle_logits`` (B, 1, A) switches to the OBB path (rotated decode + ProbIoU Fast-NMS).
    With ``sync=False`` returns the device-resident plan (rows/idx/count tensors) without any host transfer.
    ``orig_shapes`` (list of per-image (h, w[, c]), with ``img_shape`` = network input (h, w)) additionally folds the
    predictor's ``construct_result`` rescale into the gather: ``scale_boxes`` (detect/predict.py:120) or, for OBB,
    ``regularize_rboxes`` + ``scale_boxes(xywh=True)`` (obb/predict.py:59-60); rows stay cx,cy,w,h,conf,cls,angle.
    """
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    desc, keep, anchors = engine.head_desc(levels, strides, nc, reg_max)
    lv0 = keep[0]
    b = lv0.shape[0]
    rotated = angle_logits is not None
    multi_label = bool(multi_label) and nc > 1
    conf_t = _cabi.round_to_dtype(float(conf_thres), lv0.dtype)
    if rotated:
        rule, iou_eff = _cabi.RULE_FAST_PROBIOU, _cabi.f32_round(float(iou_thres))
        angle_logits = angle_logits.to(lv0.dtype).reshape(b, anchors).contiguous()
    else:
        rule, iou_eff = _cabi.RULE_GREEDY, _greedy_threshold(iou_thres)
    riders = None
    if mask_coeffs is not None or kpt_logits is not None:
        if rotated or (mask_coeffs is not None and kpt_logits is not None):
            raise ValueError("one rider at a time, and none with the rotated path")
        if kpt_logits is not None:
            if kpt_shape is None:
                raise ValueError("kpt_logits needs kpt_shape")
            riders, rider_keep = engine.riders_desc(kpt_logits.reshape(b, -1, anchors), b, anchors, lv0.dtype,
                                                    _cabi.RIDER_KEYPOINTS, int(kpt_shape[1]))
        else:
            riders, rider_keep = engine.riders_desc(mask_coeffs, b, anchors, lv0.dtype, _cabi.RIDER_RAW)
    extra = riders.channels if riders is not None else (1 if rotated else 0)
    plan = engine.make_plan(lv0.device, b, anchors, nc, extra, conf_t, iou_eff, max_det, max_nms,
                            0.0 if agnostic else float(max_wh), multi_label, rule, classes,
                            with_scale=orig_shapes is not None, cached=sync)
    if orig_shapes is not None and b:
        if img_shape is None:
            img_shape = (lv0.shape[2] * int(strides[0]), lv0.shape[3] * int(strides[0]))
        engine.set_transforms(plan, img_shape, orig_shapes, ratio_pads)
    if b and riders is not None:
        engine.run_from_head_riders(desc, riders, plan, lv0.device)
    elif b:
        engine.run_from_head(desc, angle_logits, True, plan, lv0.device)
    else:
        plan.count.zero_()
    if not sync:
        return plan
    return engine.split_results(plan, return_idxs)
