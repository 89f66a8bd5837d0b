"""Rebind the reference's own symbols to the B200 implementations.

    import ultralytics_pro_b200.patch as ypb_patch
    ypb_patch.install()        # after `import ultralytics`
    ...
    ypb_patch.uninstall()

What is rebound (reference paths relative to ultralytics/):
  utils/nms.py:13          non_max_suppression   - on the module object (seen by models/yolo/detect/predict.py:54 and
                                                   detect/val.py:115, which call `nms.non_max_suppression`) and on
                                                   nn/autobackend.py:22, which imported the function by name
  utils/nms.py:169         TorchNMS.nms / fast_nms / batched_nms - patched as static methods ON the class object, because
                                                   models/yolo/obb/val.py:14 and engine/exporter.py:122 hold the class by name
  nn/modules/head.py:151   Detect._inference (and the byte-identical copies MAFDetect :340, IDetect :535, DDetect :724);
                                                   Segment/Pose/OBB/World/YOLOE/v10 inherit it.  With lazy_decode (default) a
                                                   plain Detect head returns a lazily decoded tensor (lazy.LazyDecoded) and the
                                                   patched non_max_suppression runs the FUSED head->NMS kernels on it - the
                                                   reference's own call sites (predictor.py:335-336, detect/predict.py:54,
                                                   detect/val.py:115) reach the fused path without any change
  nn/modules/head.py:184   Detect.decode_bboxes (same four classes), head.py:1040 OBB.decode_bboxes (OBB, MAFOBB, IOBB, DOBB)
  nn/modules/block.py:250  DFL.forward - with decode_bboxes the whole of YOLOEDetect.forward_lrpc's decode (head.py:1777-1813)
  utils/ops.py:102,152,562,598,621  scale_boxes / clip_boxes / scale_coords / clip_coords / regularize_rboxes - on the module
                                                   object (every caller goes through `ops.<name>`: detect/predict.py:120,
                                                   obb/predict.py:59-60, pose/predict.py:75, detect/val.py:422, engine/results.py:341)
CUDA tensors take the kernels; anything else (CPU tensors, training mode, export) is handed to the original, untouched
reference function - that is the reference running, not a fallback of this library.
"""
from __future__ import annotations

import importlib
import sys

_saved: dict = {}

_HEAD_CLASSES = ("Detect", "MAFDetect", "IDetect", "DDetect")
_POSE_CLASSES = ("Pose", "MAFPose", "IPose", "DPose")
_OBB_CLASSES = ("OBB", "MAFOBB", "IOBB", "DOBB")


def _wrap_nms(ref_fn, ours):
    def non_max_suppression(prediction, *args, **kwargs):
        p = prediction[0] if isinstance(prediction, (list, tuple)) else prediction
        if getattr(p, "is_cuda", False):
            return ours(prediction, *args, **kwargs)
        return ref_fn(prediction, *args, **kwargs)

    non_max_suppression.__wrapped__ = ref_fn
    non_max_suppression.__doc__ = ref_fn.__doc__
    return non_max_suppression


def _kernel_ok(self) -> bool:
    """Configurations the kernels are built for; anything else runs the saved reference function."""
    import torch

    return getattr(self, "reg_max", 16) == 16 and not getattr(self, "export", False) and not torch.is_grad_enabled()


def _wrap_inference(ref_fn, ours):
    def _inference(self, x):
        if x[0].is_cuda and _kernel_ok(self) and not x[0].requires_grad:
            return ours(self, x)
        return ref_fn(self, x)

    _inference.__wrapped__ = ref_fn
    return _inference


def _wrap_decode_bboxes(ref_fn, ours):
    def decode_bboxes(self, bboxes, anchors, *args, **kwargs):
        # the tflite/edgetpu export branch of _inference passes pre-normalised tensors and is never taken here (export)
        if getattr(bboxes, "is_cuda", False) and bboxes.dim() == 3 and bboxes.shape[1] == 4 and _kernel_ok(self) \
                and not bboxes.requires_grad:
            return ours(self, bboxes, anchors, *args, **kwargs)
        return ref_fn(self, bboxes, anchors, *args, **kwargs)

    decode_bboxes.__wrapped__ = ref_fn
    return decode_bboxes


def _wrap_dfl(ref_fn, ours):
    def forward(self, x):
        import torch

        if getattr(x, "is_cuda", False) and x.dim() == 3 and self.c1 == 16 and x.shape[1] == 64 and not torch.is_grad_enabled() \
                and not x.requires_grad:
            return ours(self, x)
        return ref_fn(self, x)

    forward.__wrapped__ = ref_fn
    return forward


def _wrap_kpts(ref_fn, ours):
    def kpts_decode(self, bs, kpts):
        if kpts.is_cuda and not getattr(self, "export", False):
            return ours(self, bs, kpts)
        return ref_fn(self, bs, kpts)

    kpts_decode.__wrapped__ = ref_fn
    return kpts_decode


def _wrap_static(ref_fn, ours):
    def f(boxes, *args, **kwargs):
        if getattr(boxes, "is_cuda", False):
            return ours(boxes, *args, **kwargs)
        return ref_fn(boxes, *args, **kwargs)

    f.__wrapped__ = ref_fn
    return staticmethod(f)


_OPS_NAMES = ("scale_boxes", "clip_boxes", "scale_coords", "clip_coords", "regularize_rboxes")


def _wrap_ops(ref_fn, ours, tensor_arg: int):
    def f(*args, **kwargs):
        t = args[tensor_arg] if len(args) > tensor_arg else None
        if t is None:
            t = kwargs.get("boxes", kwargs.get("coords", kwargs.get("rboxes")))
        import torch

        if isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() <= 2 + (ref_fn.__name__.endswith("coords")) \
                and (t.numel() == 0 or t.stride(-1) == 1):
            return ours(*args, **kwargs)
        return ref_fn(*args, **kwargs)

    f.__wrapped__ = ref_fn
    f.__name__ = ref_fn.__name__
    f.__doc__ = ref_fn.__doc__
    return f


def _wrap_masks(ref_fn, ours):
    def f(protos, masks_in, bboxes, shape, *args, **kwargs):
        if getattr(protos, "is_cuda", False) and getattr(masks_in, "is_cuda", False) and protos.dim() == 3:
            return ours(protos, masks_in, bboxes, shape, *args, **kwargs)
        return ref_fn(protos, masks_in, bboxes, shape, *args, **kwargs)

    f.__wrapped__ = ref_fn
    f.__name__ = ref_fn.__name__
    return f


def _wrap_match(ref_fn, ours):
    def match_predictions(self, pred_classes, true_classes, iou, use_scipy=False):
        if getattr(iou, "is_cuda", False) and not use_scipy and len(self.iouv) <= 16:
            return ours(self, pred_classes, true_classes, iou)
        return ref_fn(self, pred_classes, true_classes, iou, use_scipy)

    match_predictions.__wrapped__ = ref_fn
    return match_predictions


def _wrap_process_batch(ref_fn, ours):
    def _process_batch(self, preds, batch):
        if getattr(preds["bboxes"], "is_cuda", False) and len(self.iouv) <= 16:
            return ours(self, preds, batch)
        return ref_fn(self, preds, batch)

    _process_batch.__wrapped__ = ref_fn
    return _process_batch


def install(lazy_decode: bool = True) -> list:
    """Patch the already-importable `ultralytics` package in place; returns the list of rebound symbols.

    lazy_decode: ``Detect._inference`` of a plain Detect head returns ``lazy.LazyDecoded`` so that the reference's
    ``_inference`` -> ``non_max_suppression`` call sites take the fused kernels (see ``lazy.py``); False keeps the two
    independent kernels (dense decode, NMS from the dense tensor)."""
    from . import head as our_head
    from . import lazy
    from . import nms as our_nms

    lazy.ENABLED = bool(lazy_decode)
    if _saved:
        return sorted(_saved)
    done = []
    ref_nms = importlib.import_module("ultralytics.utils.nms")
    _saved["ultralytics.utils.nms.non_max_suppression"] = (ref_nms, "non_max_suppression", ref_nms.non_max_suppression)
    wrapped = _wrap_nms(ref_nms.non_max_suppression, our_nms.non_max_suppression)
    ref_nms.non_max_suppression = wrapped
    done.append("ultralytics.utils.nms.non_max_suppression")
    ab = sys.modules.get("ultralytics.nn.autobackend")
    if ab is not None and hasattr(ab, "non_max_suppression"):
        _saved["ultralytics.nn.autobackend.non_max_suppression"] = (ab, "non_max_suppression", ab.non_max_suppression)
        ab.non_max_suppression = wrapped
        done.append("ultralytics.nn.autobackend.non_max_suppression")
    cls = ref_nms.TorchNMS
    for name in ("nms", "fast_nms", "batched_nms"):
        orig = cls.__dict__[name]
        _saved[f"ultralytics.utils.nms.TorchNMS.{name}"] = (cls, name, orig)
        setattr(cls, name, _wrap_static(getattr(cls, name), getattr(our_nms.TorchNMS, name)))
        done.append(f"ultralytics.utils.nms.TorchNMS.{name}")
    from . import ops as our_ops

    ref_ops = importlib.import_module("ultralytics.utils.ops")
    for name in _OPS_NAMES:
        orig = getattr(ref_ops, name)
        _saved[f"ultralytics.utils.ops.{name}"] = (ref_ops, name, orig)
        setattr(ref_ops, name, _wrap_ops(orig, getattr(our_ops, name), 0 if name in ("clip_boxes", "clip_coords", "regularize_rboxes") else 1))
        done.append(f"ultralytics.utils.ops.{name}")
    for name in ("process_mask", "process_mask_native"):  # utils/ops.py:489, :516 (segment/predict.py:101-103, segment/val.py:76)
        orig = getattr(ref_ops, name)
        _saved[f"ultralytics.utils.ops.{name}"] = (ref_ops, name, orig)
        setattr(ref_ops, name, _wrap_masks(orig, getattr(our_ops, name)))
        done.append(f"ultralytics.utils.ops.{name}")
    try:  # engine/validator.py:267, models/yolo/detect/val.py:274
        from . import val as our_val

        ref_validator = importlib.import_module("ultralytics.engine.validator")
        ref_detval = importlib.import_module("ultralytics.models.yolo.detect.val")
        bv, dv = ref_validator.BaseValidator, ref_detval.DetectionValidator
        _saved["ultralytics.engine.validator.BaseValidator.match_predictions"] = (bv, "match_predictions", bv.__dict__["match_predictions"])
        bv.match_predictions = _wrap_match(bv.__dict__["match_predictions"], our_val.match_predictions)
        done.append("ultralytics.engine.validator.BaseValidator.match_predictions")
        _saved["ultralytics.models.yolo.detect.val.DetectionValidator._process_batch"] = (dv, "_process_batch", dv.__dict__["_process_batch"])
        dv._process_batch = _wrap_process_batch(dv.__dict__["_process_batch"], our_val.process_batch)
        done.append("ultralytics.models.yolo.detect.val.DetectionValidator._process_batch")
    except Exception:  # validator stack not importable in this environment: matching stays with the reference
        pass
    # fast_nms resolves iou_func by __name__, so the reference's own box_iou / batch_probiou callables are recognised
    try:
        ref_head = importlib.import_module("ultralytics.nn.modules.head")
    except Exception:  # optional third-party imports of the modules zoo missing: decode stays with the reference
        ref_head = None
    if ref_head is not None:
        for cname in _HEAD_CLASSES:
            c = getattr(ref_head, cname, None)
            if c is None or "_inference" not in c.__dict__:
                continue
            _saved[f"ultralytics.nn.modules.head.{cname}._inference"] = (c, "_inference", c.__dict__["_inference"])
            c._inference = _wrap_inference(c.__dict__["_inference"], our_head.detect_inference)
            done.append(f"ultralytics.nn.modules.head.{cname}._inference")
        for cname in _HEAD_CLASSES:  # head.py:184-191 (copies :374, :569, :757)
            c = getattr(ref_head, cname, None)
            if c is None or "decode_bboxes" not in c.__dict__:
                continue
            _saved[f"ultralytics.nn.modules.head.{cname}.decode_bboxes"] = (c, "decode_bboxes", c.__dict__["decode_bboxes"])
            c.decode_bboxes = _wrap_decode_bboxes(c.__dict__["decode_bboxes"], our_head.detect_decode_bboxes)
            done.append(f"ultralytics.nn.modules.head.{cname}.decode_bboxes")
        for cname in _OBB_CLASSES:  # head.py:1040, :1094, :1148, :1203
            c = getattr(ref_head, cname, None)
            if c is None or "decode_bboxes" not in c.__dict__:
                continue
            _saved[f"ultralytics.nn.modules.head.{cname}.decode_bboxes"] = (c, "decode_bboxes", c.__dict__["decode_bboxes"])
            c.decode_bboxes = _wrap_decode_bboxes(c.__dict__["decode_bboxes"], our_head.obb_decode_bboxes)
            done.append(f"ultralytics.nn.modules.head.{cname}.decode_bboxes")
        dfl = getattr(ref_head, "DFL", None)  # head.py imports DFL from .block by name; patching the class covers both
        if dfl is not None and "forward" in dfl.__dict__:
            _saved["ultralytics.nn.modules.block.DFL.forward"] = (dfl, "forward", dfl.__dict__["forward"])
            dfl.forward = _wrap_dfl(dfl.__dict__["forward"], our_head.dfl_forward)
            done.append("ultralytics.nn.modules.block.DFL.forward")
        for cname in _HEAD_CLASSES:  # head.py:193 end2end top-k (staticmethod)
            c = getattr(ref_head, cname, None)
            if c is None or "postprocess" not in c.__dict__:
                continue
            _saved[f"ultralytics.nn.modules.head.{cname}.postprocess"] = (c, "postprocess", c.__dict__["postprocess"])
            c.postprocess = _wrap_static(c.postprocess, our_head.detect_postprocess)
            done.append(f"ultralytics.nn.modules.head.{cname}.postprocess")
        for cname in _POSE_CLASSES:  # head.py:1254, :1322, :1390, :1459
            c = getattr(ref_head, cname, None)
            if c is None or "kpts_decode" not in c.__dict__:
                continue
            _saved[f"ultralytics.nn.modules.head.{cname}.kpts_decode"] = (c, "kpts_decode", c.__dict__["kpts_decode"])
            c.kpts_decode = _wrap_kpts(c.__dict__["kpts_decode"], our_head.pose_kpts_decode)
            done.append(f"ultralytics.nn.modules.head.{cname}.kpts_decode")
    return done


def uninstall() -> None:
    from . import lazy

    for _, (obj, name, orig) in list(_saved.items()):
        setattr(obj, name, orig)
    _saved.clear()
    lazy.ENABLED = False
