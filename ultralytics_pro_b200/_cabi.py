"""ctypes binding of libyolopost_b200.so (include/yolopost_b200.h).

There is no CPU fallback: importing this module without the built library, or calling it with CPU tensors,
raises.  PyTorch is used only for device memory (caching allocator) and the current stream.
"""
from __future__ import annotations

import ctypes as C
import functools
import os
import struct

import torch

from . import build as _build

YPB_F32, YPB_F16, YPB_BF16 = 0, 1, 2
RULE_GREEDY, RULE_FAST_PROBIOU, RULE_FAST_BOXIOU = 0, 1, 2
MAX_LEVELS = 8
MAX_PEERS = 8
ABI_VERSION = 11

_DTYPES = {torch.float32: YPB_F32, torch.float16: YPB_F16, torch.bfloat16: YPB_BF16}

EXPORTS = (
    "ypb_abi_version",
    "ypb_last_error_string",
    "ypb_nms_workspace_bytes",
    "ypb_decode_dense",
    "ypb_nms_from_head",
    "ypb_nms_from_head_stage",
    "ypb_nms_from_dense",
    "ypb_nms_boxes_workspace_bytes",
    "ypb_nms_boxes",
    "ypb_selftest_sigmoid_monotone",
    "ypb_debug_set_phase_buffer",
    "ypb_scale_rows",
    "ypb_nms_from_head_riders",
    "ypb_kpts_decode",
    "ypb_process_mask",
    "ypb_match_predictions",
    "ypb_peer_wait",
    "ypb_dfl_expectation",
    "ypb_dist2bbox",
    "ypb_pairwise_iou",
    "ypb_compact_results",
    "ypb_peer_wait_copy",
    "ypb_process_mask_workspace_bytes",
)
MASK_CROP_PROTO, MASK_CROP_OUTPUT = 1, 2
RIDER_RAW, RIDER_KEYPOINTS = 0, 1
SCAN_AUTO, SCAN_LDG, SCAN_TMA = 0, 1, 2

BOXES_NONE, BOXES_XYXY, BOXES_XYWH, BOXES_XYWHR, BOXES_CLIP_ONLY, BOXES_REGULARIZE_ONLY = range(6)
SCALE_PADDING, SCALE_NORMALIZE, SCALE_COORDS_CLIP_ONLY = 1, 2, 4


class HeadDesc(C.Structure):
    _fields_ = [
        ("num_levels", C.c_int32),
        ("batch", C.c_int32),
        ("nc", C.c_int32),
        ("reg_max", C.c_int32),
        ("dtype", C.c_int32),
        ("reserved", C.c_int32),
        ("level_ptr", C.c_void_p * MAX_LEVELS),
        ("level_h", C.c_int32 * MAX_LEVELS),
        ("level_w", C.c_int32 * MAX_LEVELS),
        ("level_batch_stride", C.c_int64 * MAX_LEVELS),
        ("level_channel_stride", C.c_int64 * MAX_LEVELS),
        ("level_stride", C.c_float * MAX_LEVELS),
    ]


class DenseDesc(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("dtype", C.c_int32),
        ("batch", C.c_int32),
        ("channels", C.c_int32),
        ("anchors", C.c_int32),
        ("stride_b", C.c_int64),
        ("stride_c", C.c_int64),
        ("stride_a", C.c_int64),
        ("anchor_subset", C.c_void_p),
        ("subset_len", C.c_int32),
        ("reserved", C.c_int32),
    ]


class NmsParams(C.Structure):
    _fields_ = [
        ("conf_thres", C.c_float),
        ("iou_thres_eff", C.c_float),
        ("nc", C.c_int32),
        ("extra", C.c_int32),
        ("max_det", C.c_int32),
        ("max_nms", C.c_int32),
        ("max_wh", C.c_float),
        ("multi_label", C.c_int32),
        ("rule", C.c_int32),
        ("rows_cap", C.c_int32),
        ("class_mask", C.c_void_p),
        ("nms_box_divisor", C.c_float),
        ("nms_box_multiplier", C.c_float),
        ("boxes_xyxy", C.c_int32),
        ("pad_output", C.c_int32),
        ("conf_per_image", C.c_void_p),
        ("clean_counters", C.c_void_p),
        ("scan_kernel", C.c_int32),
        ("reserved2", C.c_int32),
    ]


class NmsOut(C.Structure):
    _fields_ = [
        ("rows", C.c_void_p),
        ("idx", C.c_void_p),
        ("count", C.c_void_p),
        ("cand_count", C.c_void_p),
        ("scale_xforms", C.c_void_p),
        ("scale_padding", C.c_int32),
        ("num_peers", C.c_int32),
        ("my_rank", C.c_int32),
        ("peer_depth", C.c_int32),
        ("peer_rows", C.c_void_p * MAX_PEERS),
        ("peer_count", C.c_void_p * MAX_PEERS),
        ("peer_flag", C.c_void_p * MAX_PEERS),
        ("peer_state", C.c_void_p),
        ("peer_ack", C.c_void_p),
        ("peer_entry_stride", C.c_int64),
        ("count_host", C.c_void_p),
    ]


class RidersDesc(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("channels", C.c_int32),
        ("kind", C.c_int32),
        ("kpt_ndim", C.c_int32),
        ("reserved", C.c_int32),
        ("stride_b", C.c_int64),
        ("stride_c", C.c_int64),
    ]


class ProtosDesc(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("dtype", C.c_int32),
        ("channels", C.c_int32),
        ("mh", C.c_int32),
        ("mw", C.c_int32),
        ("stride_b", C.c_int64),
        ("stride_c", C.c_int64),
    ]


class ScaleXform(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("gain", "pad_x", "pad_y", "img_w", "img_h", "cpad_x", "cpad_y", "reserved")]


_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building in-tree if needed) the shared library; raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if not os.path.exists(path) or (not _build.is_current() and _can_build()):
        path = _build.build_library()
    lib = C.CDLL(path)
    missing = [n for n in EXPORTS if not hasattr(lib, n)]
    if missing:
        raise RuntimeError(f"{path} does not export {missing}")
    lib.ypb_abi_version.restype = C.c_int
    lib.ypb_last_error_string.restype = C.c_char_p
    lib.ypb_nms_workspace_bytes.restype = C.c_size_t
    lib.ypb_nms_workspace_bytes.argtypes = [C.c_int32] * 6
    lib.ypb_decode_dense.restype = C.c_int
    lib.ypb_decode_dense.argtypes = [C.POINTER(HeadDesc), C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                     C.c_int32, C.c_int64, C.c_int64, C.c_void_p]
    lib.ypb_nms_from_head.restype = C.c_int
    lib.ypb_nms_from_head.argtypes = [C.POINTER(HeadDesc), C.c_void_p, C.c_int32, C.c_int32, C.POINTER(NmsParams),
                                      C.POINTER(NmsOut), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.ypb_nms_from_head_stage.restype = C.c_int
    lib.ypb_nms_from_head_stage.argtypes = lib.ypb_nms_from_head.argtypes + [C.c_int32]
    lib.ypb_nms_from_dense.restype = C.c_int
    lib.ypb_nms_from_dense.argtypes = [C.POINTER(DenseDesc), C.POINTER(NmsParams), C.POINTER(NmsOut), C.c_void_p,
                                       C.c_size_t, C.c_void_p]
    lib.ypb_nms_boxes_workspace_bytes.restype = C.c_size_t
    lib.ypb_nms_boxes_workspace_bytes.argtypes = [C.c_int32]
    lib.ypb_nms_boxes.restype = C.c_int
    lib.ypb_nms_boxes.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.ypb_debug_set_phase_buffer.restype = None
    lib.ypb_debug_set_phase_buffer.argtypes = [C.c_void_p]
    lib.ypb_selftest_sigmoid_monotone.restype = C.c_int
    lib.ypb_selftest_sigmoid_monotone.argtypes = [C.c_int32, C.c_void_p, C.c_void_p]
    lib.ypb_scale_rows.restype = C.c_int
    lib.ypb_scale_rows.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.POINTER(ScaleXform), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                   C.c_int64, C.c_int32, C.c_int32, C.c_void_p]
    lib.ypb_nms_from_head_riders.restype = C.c_int
    lib.ypb_nms_from_head_riders.argtypes = [C.POINTER(HeadDesc), C.POINTER(RidersDesc), C.c_int32, C.POINTER(NmsParams),
                                             C.POINTER(NmsOut), C.c_void_p, C.c_size_t, C.c_void_p]
    lib.ypb_kpts_decode.restype = C.c_int
    lib.ypb_kpts_decode.argtypes = [C.POINTER(HeadDesc), C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p]
    lib.ypb_process_mask.restype = C.c_int
    lib.ypb_process_mask.argtypes = [C.POINTER(ProtosDesc), C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                     C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_void_p]
    lib.ypb_process_mask_workspace_bytes.restype = C.c_size_t
    lib.ypb_process_mask_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
    lib.ypb_match_predictions.restype = C.c_int
    lib.ypb_match_predictions.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.POINTER(C.c_float), C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_size_t, C.c_void_p]
    lib.ypb_peer_wait.restype = C.c_int
    lib.ypb_peer_wait.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_int32,
                                  C.c_void_p, C.c_void_p]
    lib.ypb_dfl_expectation.restype = C.c_int
    lib.ypb_dfl_expectation.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64,
                                        C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.ypb_dist2bbox.restype = C.c_int
    lib.ypb_dist2bbox.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64,
                                  C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.ypb_pairwise_iou.restype = C.c_int
    lib.ypb_pairwise_iou.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.ypb_compact_results.restype = C.c_int
    lib.ypb_compact_results.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    lib.ypb_peer_wait_copy.restype = C.c_int
    lib.ypb_peer_wait_copy.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    if lib.ypb_abi_version() != ABI_VERSION:
        raise RuntimeError(f"{path}: ABI version {lib.ypb_abi_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def _can_build() -> bool:
    try:
        _build._nvcc()
        return True
    except RuntimeError:
        return False


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().ypb_last_error_string().decode(errors="replace")
        raise RuntimeError(f"{what} failed ({rc}): {msg}")


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise TypeError(f"unsupported dtype {dt}; the kernels take float32, float16 and bfloat16") from None


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor, got {t.device}; this path has no CPU fallback")


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def f32_round(v: float) -> float:
    """Nearest float32 to the Python float (what torch does to a scalar operand of an fp32 tensor op)."""
    return struct.unpack("f", struct.pack("f", v))[0]


def largest_f32_not_above(v: float) -> float:
    """Largest float32 <= v: `x > v` evaluated in double (torchvision's CPU nms) for fp32 x <=> `x > this` in fp32."""
    f = f32_round(v)
    if f <= v:
        return f
    bits = struct.unpack("I", struct.pack("f", f))[0]
    bits = bits - 1 if f > 0 else bits + 1
    return struct.unpack("f", struct.pack("I", bits))[0]


@functools.lru_cache(maxsize=256)
def round_to_dtype(v: float, dt: torch.dtype) -> float:
    """Value of the Python scalar after torch casts it to `dt` for a tensor-scalar comparison (nms.py:76,115,121)."""
    return float(torch.tensor(v, dtype=torch.float64).to(dt).to(torch.float64))
