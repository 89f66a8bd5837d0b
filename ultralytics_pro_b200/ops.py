"""Drop-ins for the result-side helpers of ``ultralytics/utils/ops.py`` that run right after NMS:

  scale_boxes (ops.py:102-135)   clip_boxes (ops.py:152-177)    regularize_rboxes (ops.py:621-636)
  scale_coords (ops.py:562-595)  clip_coords (ops.py:598-618)

Same names, arguments and in-place behaviour as the reference; each call is ONE launch of ``scale_rows_kernel``
(libyolopost_b200) instead of the reference's 9-13 ATen launches.  ``scale_results`` applies the same arithmetic to the
kept rows of a whole batch (per-image transform and count, still one launch, no host synchronisation).
Arithmetic follows the reference's CPU results bit for bit: fp32, every step separately rounded, IEEE division.
There is no CPU path: CPU tensors raise (``patch.install`` leaves those with the reference).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi

_f32 = _cabi.f32_round
TWO_STEP_MIN_BYTES = 1 << 20  # process_mask results of at least this size take the memset + work-list form


def letterbox_transform(img1_shape, img0_shape, ratio_pad=None) -> _cabi.ScaleXform:
    """Host scalars of ops.py:120-127 (boxes) and ops.py:580-587 (coords) for one image, rounded to fp32."""
    h0, w0 = img0_shape[:2]
    if ratio_pad is None:
        h1, w1 = img1_shape[:2]
        gain = min(h1 / h0, w1 / w0)
        cpad_x, cpad_y = (w1 - w0 * gain) / 2, (h1 - h0 * gain) / 2
        pad_x, pad_y = round(cpad_x - 0.1), round(cpad_y - 0.1)
    else:
        gain = ratio_pad[0][0]
        pad_x, pad_y = ratio_pad[1]
        cpad_x, cpad_y = pad_x, pad_y
    x = _cabi.ScaleXform()
    x.gain, x.pad_x, x.pad_y, x.img_w, x.img_h = _f32(gain), _f32(pad_x), _f32(pad_y), _f32(w0), _f32(h0)
    x.cpad_x, x.cpad_y = _f32(cpad_x), _f32(cpad_y)
    return x


def _rows2d(t: torch.Tensor, min_cols: int, what: str) -> torch.Tensor:
    _cabi.require_cuda(t, what)
    if t.dtype != torch.float32:
        raise TypeError(f"{what}: expected float32 result rows (nms.py:116 promotes them), got {t.dtype}")
    if t.shape[-1] < min_cols:
        raise ValueError(f"{what}: last dimension {t.shape[-1]} < {min_cols}")
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.dim() > 2:
        t = t.view(-1, t.shape[-1])  # raises for layouts that are not a uniform stack of rows
    if t.shape[0] and t.stride(1) != 1:
        raise ValueError(f"{what}: the coordinates of a row must be contiguous")
    return t


def _launch(rows2d, xf, box_mode, flags=0, angle_col=0, coords=None, coord_row_stride=0, nk=0, ndim=0):
    lib = _cabi.load()
    base = rows2d if rows2d is not None else coords
    n = base.shape[0]
    if n == 0:
        return
    rc = lib.ypb_scale_rows(rows2d.data_ptr() if rows2d is not None else None, 0, rows2d.stride(0) if rows2d is not None else 0,
                            1, n, None, None, C.byref(xf), box_mode, flags, angle_col,
                            coords.data_ptr() if coords is not None else None, 0, coord_row_stride, nk, ndim,
                            _cabi.stream_ptr(base.device))
    _cabi.check(rc, "ypb_scale_rows")


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding: bool = True, xywh: bool = False):
    """Rescale boxes (N, 4) from ``img1_shape`` (network input) to ``img0_shape`` in place; mirror of ops.py:102-135."""
    r = _rows2d(boxes, 4, "scale_boxes")
    xf = letterbox_transform(img1_shape, img0_shape, ratio_pad)
    _launch(r, xf, _cabi.BOXES_XYWH if xywh else _cabi.BOXES_XYXY, _cabi.SCALE_PADDING if padding else 0)
    return boxes


def clip_boxes(boxes, shape):
    """Clamp xyxy boxes to the image in place; mirror of ops.py:152-177."""
    r = _rows2d(boxes, 4, "clip_boxes")
    xf = _cabi.ScaleXform()
    xf.gain, xf.img_w, xf.img_h = 1.0, _f32(shape[1]), _f32(shape[0])
    _launch(r, xf, _cabi.BOXES_CLIP_ONLY)
    return boxes


def regularize_rboxes(rboxes):
    """Angles into [0, pi/2) with w/h swapped where needed; returns a new (N, 5) tensor like ops.py:621-636."""
    _cabi.require_cuda(rboxes, "regularize_rboxes")
    out = rboxes.to(torch.float32).clone(memory_format=torch.contiguous_format)
    r = _rows2d(out, 5, "regularize_rboxes")
    xf = _cabi.ScaleXform()
    xf.gain = 1.0
    _launch(r, xf, _cabi.BOXES_REGULARIZE_ONLY, angle_col=4)
    return out


def scale_coords(img1_shape, coords, img0_shape, ratio_pad=None, normalize: bool = False, padding: bool = True):
    """Rescale points (..., 2|3) in place (x, y are the first two of the last dimension); mirror of ops.py:562-595."""
    _cabi.require_cuda(coords, "scale_coords")
    if coords.dtype != torch.float32:
        raise TypeError(f"scale_coords: expected float32, got {coords.dtype}")
    ndim = coords.shape[-1]
    pts = coords.view(-1, ndim)
    if pts.shape[0] and pts.stride(1) != 1:
        raise ValueError("scale_coords: x and y of a point must be adjacent")
    xf = letterbox_transform(img1_shape, img0_shape, ratio_pad)
    flags = (_cabi.SCALE_PADDING if padding else 0) | (_cabi.SCALE_NORMALIZE if normalize else 0)
    _launch(None, xf, _cabi.BOXES_NONE, flags, coords=pts, coord_row_stride=pts.stride(0) if pts.shape[0] else ndim, nk=1, ndim=ndim)
    return coords


def clip_coords(coords, shape):
    """Clamp points to the image in place; mirror of ops.py:598-618."""
    _cabi.require_cuda(coords, "clip_coords")
    if coords.dtype != torch.float32:
        raise TypeError(f"clip_coords: expected float32, got {coords.dtype}")
    ndim = coords.shape[-1]
    pts = coords.view(-1, ndim)
    xf = _cabi.ScaleXform()
    xf.gain, xf.img_w, xf.img_h = 1.0, _f32(shape[1]), _f32(shape[0])
    _launch(None, xf, _cabi.BOXES_NONE, _cabi.SCALE_COORDS_CLIP_ONLY, coords=pts,
            coord_row_stride=pts.stride(0) if pts.shape[0] else ndim, nk=1, ndim=ndim)
    return coords


# ----------------------------------------------------------------------------------------------------------------
# batched form: the kept rows of a whole batch in one launch
# ----------------------------------------------------------------------------------------------------------------
_xform_cache: dict = {}


def transforms_tensor(img1_shape, orig_shapes, ratio_pads=None, device=None) -> torch.Tensor:
    """(B, 8) fp32 device array of ``ypb_scale_xform`` for a batch; cached per (shapes, device)."""
    key = (tuple(img1_shape[:2]), tuple(tuple(s[:2]) for s in orig_shapes),
           None if ratio_pads is None else tuple((tuple(r[0]), tuple(r[1])) for r in ratio_pads), str(device))
    t = _xform_cache.get(key)
    if t is None:
        rows = []
        for i, s0 in enumerate(orig_shapes):
            x = letterbox_transform(img1_shape, s0, None if ratio_pads is None else ratio_pads[i])
            rows.append([x.gain, x.pad_x, x.pad_y, x.img_w, x.img_h, x.cpad_x, x.cpad_y, 0.0])
        t = torch.tensor(rows, dtype=torch.float32).pin_memory().to(device, non_blocking=True)
        if len(_xform_cache) > 256:
            _xform_cache.clear()
        _xform_cache[key] = t
    return t


def scale_results(rows: torch.Tensor, count: torch.Tensor | None, img1_shape, orig_shapes, ratio_pads=None,
                  padding: bool = True, rotated: bool = False, kpt_shape=None, normalize_kpts: bool = False):
    """In-place ``construct_result`` scaling (detect/predict.py:120, obb/predict.py:59-60, pose/predict.py:73-75) of the
    padded result rows ``(B, max_det, 6+extra)`` of a batch, honouring the per-image kept ``count`` (device int32) -
    one launch, nothing read back.  ``rotated``: rows are cx,cy,w,h,conf,cls,angle -> regularize + xywh scaling;
    ``kpt_shape``: (nk, ndim) keypoints in columns 6.. are scaled like ``scale_coords``."""
    _cabi.require_cuda(rows, "scale_results")
    if rows.dtype != torch.float32 or rows.dim() != 3 or rows.stride(2) != 1:
        raise ValueError("scale_results: rows must be a (B, max_det, cols) float32 tensor with contiguous columns")
    b, m, cols = rows.shape
    if len(orig_shapes) != b:
        raise ValueError(f"{len(orig_shapes)} original shapes for a batch of {b}")
    xf = transforms_tensor(img1_shape, orig_shapes, ratio_pads, rows.device)
    mode = _cabi.BOXES_XYWHR if rotated else _cabi.BOXES_XYXY
    nk = ndim = 0
    coords_ptr = None
    if kpt_shape is not None:
        nk, ndim = int(kpt_shape[0]), int(kpt_shape[1])
        if cols < 6 + nk * ndim:
            raise ValueError(f"rows have {cols} columns, keypoints need {6 + nk * ndim}")
        coords_ptr = rows.data_ptr() + 6 * 4
    flags = (_cabi.SCALE_PADDING if padding else 0) | (_cabi.SCALE_NORMALIZE if normalize_kpts else 0)
    lib = _cabi.load()
    rc = lib.ypb_scale_rows(rows.data_ptr(), rows.stride(0), rows.stride(1), b, m,
                            count.data_ptr() if count is not None else None, xf.data_ptr(), None, mode, flags,
                            cols - 1 if rotated else 0, coords_ptr, rows.stride(0), rows.stride(1), nk, ndim,
                            _cabi.stream_ptr(rows.device))
    _cabi.check(rc, "ypb_scale_rows")
    return rows


# ----------------------------------------------------------------------------------------------------------------
# Segment masks (utils/ops.py:489-559)
# ----------------------------------------------------------------------------------------------------------------
def _protos_desc(protos: torch.Tensor):
    _cabi.require_cuda(protos, "process_mask")
    if protos.dim() == 3:
        protos = protos.unsqueeze(0)
    if protos.dim() != 4:
        raise ValueError(f"protos must be (C, mh, mw) or (B, C, mh, mw), got {tuple(protos.shape)}")
    if protos.stride(3) != 1 or protos.stride(2) != protos.shape[3]:
        protos = protos.contiguous()
    d = _cabi.ProtosDesc()
    d.ptr, d.dtype = protos.data_ptr(), _cabi.dtype_code(protos.dtype)
    _, d.channels, d.mh, d.mw = protos.shape
    d.stride_b, d.stride_c = protos.stride(0), protos.stride(1)
    return d, protos


def _f32_rows(t: torch.Tensor, cols: int, what: str) -> torch.Tensor:
    _cabi.require_cuda(t, what)
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2 or t.shape[1] != cols:
        raise ValueError(f"{what}: expected (n, {cols}), got {tuple(t.shape)}")
    if t.shape[0] and t.stride(1) != 1:
        t = t.contiguous()
    return t


def _scale_masks_window(mh: int, mw: int, shape, padding: bool = True):
    """Rows / columns of the prototype grid that scale_masks resizes (ops.py:544-559)."""
    gain = min(mh / shape[0], mw / shape[1])
    pad_w, pad_h = mw - shape[1] * gain, mh - shape[0] * gain
    if padding:
        pad_w /= 2
        pad_h /= 2
    top, left = (round(pad_h - 0.1), round(pad_w - 0.1)) if padding else (0, 0)
    bottom, right = mh - round(pad_h + 0.1), mw - round(pad_w + 0.1)
    return top, left, bottom - top, right - left


def _run_masks(pd, coeffs, cis, crs, boxes, bis, brs, offsets, batch, total, out_hw, window, crop_mode, ratios, device):
    out = torch.empty((total, out_hw[0], out_hw[1]), dtype=torch.uint8, device=device)
    if total:
        lib = _cabi.load()
        # two-step form (memset + work list + compute only the tiles that can see their box) when the result is large enough
        # for the zero-fill to matter; TWO_STEP_MIN_BYTES = 0 / huge forces one or the other (tests cover both)
        ws, ws_bytes = None, 0
        if out.numel() >= TWO_STEP_MIN_BYTES:
            from . import engine

            ws_bytes = lib.ypb_process_mask_workspace_bytes(total, out_hw[0], out_hw[1])
            ws = engine._scratch(device, ws_bytes)
        rc = lib.ypb_process_mask(C.byref(pd), coeffs.data_ptr(), cis, crs, boxes.data_ptr(), bis, brs,
                                  offsets.data_ptr() if offsets is not None else None, batch, total,
                                  out_hw[0], out_hw[1], window[0], window[1], window[2], window[3], crop_mode,
                                  ratios[0], ratios[1], out.data_ptr(), ws.data_ptr() if ws is not None else None, ws_bytes,
                                  _cabi.stream_ptr(device))
        _cabi.check(rc, "ypb_process_mask")
    return out


def process_mask(protos, masks_in, bboxes, shape, upsample: bool = False):
    """Mirror of ops.py:489-513: (n, H, W) uint8 masks of one image - ``masks_in @ protos`` cropped to the boxes at
    prototype resolution, bilinearly upsampled to ``shape`` when ``upsample``, thresholded at 0.  One fused kernel;
    follows the reference's CUDA / n >= 50 ``crop_mask`` branch (ops.py:482-486)."""
    pd, keep = _protos_desc(protos)
    c, mh, mw = pd.channels, pd.mh, pd.mw
    co, bx = _f32_rows(masks_in, c, "process_mask coefficients"), _f32_rows(bboxes, 4, "process_mask boxes")
    n = co.shape[0]
    out_hw = (int(shape[0]), int(shape[1])) if upsample else (mh, mw)
    ratios = (_f32(mw / shape[1]), _f32(mh / shape[0]))
    return _run_masks(pd, co, 0, co.stride(0) if n else c, bx, 0, bx.stride(0) if n else 4, None, 1, n, out_hw,
                      (0, 0, mh, mw), _cabi.MASK_CROP_PROTO, ratios, keep.device)


def process_mask_native(protos, masks_in, bboxes, shape):
    """Mirror of ops.py:516-541: masks resized to ``shape`` with the letterbox padding removed (scale_masks, ops.py:544-559),
    then cropped to the boxes at that resolution, thresholded at 0."""
    pd, keep = _protos_desc(protos)
    c, mh, mw = pd.channels, pd.mh, pd.mw
    co, bx = _f32_rows(masks_in, c, "process_mask_native coefficients"), _f32_rows(bboxes, 4, "process_mask_native boxes")
    n = co.shape[0]
    win = _scale_masks_window(mh, mw, shape)
    return _run_masks(pd, co, 0, co.stride(0) if n else c, bx, 0, bx.stride(0) if n else 4, None, 1, n,
                      (int(shape[0]), int(shape[1])), win, _cabi.MASK_CROP_OUTPUT, (1.0, 1.0), keep.device)


def process_masks_batched(protos, rows, counts, shape, upsample: bool = True):
    """``process_mask`` for every image of a batch in ONE launch (segment/predict.py:84-103 loops over images).

    protos (B, C, mh, mw); rows (B, max_det, 6+C) padded NMS result rows (boxes in columns 0..3 in network-input pixels,
    coefficients in 6..); counts: per-image kept counts on the HOST (``engine.fetch_counts``).  Returns the list of
    (n_i, H, W) uint8 views of one packed buffer."""
    pd, keep = _protos_desc(protos)
    c, mh, mw = pd.channels, pd.mh, pd.mw
    _cabi.require_cuda(rows, "process_masks_batched")
    if rows.dtype != torch.float32 or rows.dim() != 3 or rows.shape[2] != 6 + c or rows.stride(2) != 1:
        raise ValueError(f"rows must be (B, max_det, {6 + c}) float32 with contiguous columns, got {tuple(rows.shape)}")
    b = rows.shape[0]
    if len(counts) != b or keep.shape[0] != b:
        raise ValueError("counts / protos do not match the batch")
    offs = [0]
    for n in counts:
        offs.append(offs[-1] + int(n))
    total = offs[-1]
    offsets = torch.tensor(offs, dtype=torch.int32, device=rows.device)  # small pageable H2D: staged by the driver, no pinned alloc
    out_hw = (int(shape[0]), int(shape[1])) if upsample else (mh, mw)
    ratios = (_f32(mw / shape[1]), _f32(mh / shape[0]))
    coeffs = rows[:, :, 6:]
    packed = _run_masks(pd, coeffs, rows.stride(0), rows.stride(1), rows, rows.stride(0), rows.stride(1), offsets, b, total,
                        out_hw, (0, 0, mh, mw), _cabi.MASK_CROP_PROTO, ratios, rows.device)
    return [packed[offs[i]:offs[i + 1]] for i in range(b)]
