"""Side channel between the two halves of the drop-in: ``Detect._inference`` -> ``non_max_suppression``.

The reference's call sites (``engine/predictor.py:335-336``, ``models/yolo/detect/predict.py:54-65``,
``detect/val.py:115``) hand the dense ``(B, 4+nc, A)`` tensor returned by ``Detect._inference`` (head.py:151-169) straight
to ``non_max_suppression`` (utils/nms.py:13).  Decoding all A anchors into HBM only to throw > 99 % of them away in the
confidence filter is what the fused head->NMS kernels avoid - but they need the raw level tensors, which the NMS call
never sees.  ``LazyDecoded`` closes that gap without changing either signature: it IS a ``torch.Tensor`` (wrapper
subclass: shape, dtype, device and strides of the dense result, no storage) that remembers the level tensors.

  * Any torch operator applied to it (``cat``, ``permute``, indexing, ``.cpu()``, in-place writes, printing ...) reaches
    ``__torch_dispatch__``, which runs the dense decode kernel once (``materialize``) and re-dispatches on the real tensor -
    so every consumer other than our NMS sees exactly what ``Detect._inference`` would have returned.
  * The patched ``non_max_suppression`` asks ``head_record()``; when the tensor is still pending it runs
    ``postprocess_from_head`` on the recorded levels (bit-identical to decode + NMS, tests/test_gpu_parity.py) and the
    dense tensor is never written.

The level tensors must not be modified between the two calls (true for every reference call site: ``Detect.forward``
returns them as the second element of its tuple for the loss, nobody writes to them in eval mode).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch.utils._pytree import tree_map

ENABLED = False  # set by patch.install(lazy_decode=True)
STATS = {"created": 0, "materialized": 0, "fused": 0}


@dataclass
class HeadRecord:
    levels: list
    strides: tuple
    nc: int
    reg_max: int
    xyxy: bool


def _decode(rec: HeadRecord) -> torch.Tensor:
    from .head import decode_head

    return decode_head(rec.levels, rec.strides, rec.nc, rec.reg_max, xyxy=rec.xyxy)


_materializer = _decode  # tests on a CPU-only box substitute the oracle decode here


class LazyDecoded(torch.Tensor):
    """Dense ``Detect._inference`` result that is decoded on first use (see the module docstring)."""

    @staticmethod
    def __new__(cls, shape, dtype, device):
        return torch.Tensor._make_wrapper_subclass(cls, tuple(shape), dtype=dtype, device=device, requires_grad=False)

    @classmethod
    def from_head(cls, levels, strides, nc: int, reg_max: int, xyxy: bool) -> "LazyDecoded":
        lv0 = levels[0]
        anchors = sum(int(lv.shape[2]) * int(lv.shape[3]) for lv in levels)
        r = cls((int(lv0.shape[0]), 4 + int(nc), anchors), lv0.dtype, lv0.device)
        r._ypb_rec = HeadRecord(list(levels), tuple(strides), int(nc), int(reg_max), bool(xyxy))
        r._ypb_real = None
        STATS["created"] += 1
        return r

    def head_record(self):
        """The recorded head while the dense tensor has not been needed yet, else None."""
        return self._ypb_rec if self._ypb_real is None else None

    def materialize(self) -> torch.Tensor:
        if self._ypb_real is None:
            self._ypb_real = _materializer(self._ypb_rec)
            self._ypb_rec = None  # drop the references to the level tensors
            STATS["materialized"] += 1
        return self._ypb_real

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        def unwrap(t):
            return t.materialize() if isinstance(t, LazyDecoded) else t

        return func(*tree_map(unwrap, args), **tree_map(unwrap, kwargs or {}))

    __torch_function__ = torch._C._disabled_torch_function_impl

    def __repr__(self):  # avoid materialising just to print the placeholder
        state = "pending" if self._ypb_real is None else "materialized"
        return f"LazyDecoded(shape={tuple(self.shape)}, dtype={self.dtype}, device={self.device}, {state})"
