"""Host-side sequencing shared by the drop-in functions: descriptors, scratch reuse, result views.

Everything here is plumbing around the C-ABI calls; no arithmetic of the path is done in Python/PyTorch.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass

import torch

from . import _cabi

_tls = threading.local()


def _scratch(device: torch.device, nbytes: int) -> torch.Tensor:
    """Per-(thread, device, stream) scratch buffer, grown geometrically; 256-byte aligned by the caching allocator."""
    pool = getattr(_tls, "pool", None)
    if pool is None:
        pool = _tls.pool = {}
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = pool.get(key)
    if buf is None or buf.numel() < nbytes:
        with torch.inference_mode(False):
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        pool[key] = buf
    return buf


def _pinned_counts(n: int) -> torch.Tensor:
    buf = getattr(_tls, "pinned", None)
    if buf is None or buf.numel() < n:
        with torch.inference_mode(False):  # a cached buffer must stay writable outside the inference_mode it was first needed in
            buf = _tls.pinned = torch.empty(max(n, 256), dtype=torch.int32, pin_memory=True)
    return buf[:n]


_class_masks: dict = {}


def class_mask_tensor(classes, nc: int, device: torch.device):
    """Device bitmask of the `classes` filter (nms.py:62-63,127-131); cached per (classes, nc, device)."""
    if classes is None:
        return None
    if isinstance(classes, torch.Tensor):
        classes = classes.tolist()
    key = (tuple(int(c) for c in classes), nc, device.index)
    t = _class_masks.get(key)
    if t is None:
        words = [0] * ((nc + 31) // 32)
        for c in key[0]:
            if 0 <= c < nc:
                words[c >> 5] |= 1 << (c & 31)
        # uint32 words stored as their int32 bit patterns
        t = torch.tensor([w - (1 << 32) if w >= (1 << 31) else w for w in words], dtype=torch.int32, device=device)
        _class_masks[key] = t
    return t


def head_desc(levels, strides, nc: int, reg_max: int):
    """Build the ypb_head_desc for a list of (B, 4*reg_max+nc, H, W) level tensors (head.py:121-122)."""
    if not levels:
        raise ValueError("no head levels given")
    if len(levels) > _cabi.MAX_LEVELS:
        raise ValueError(f"{len(levels)} levels > {_cabi.MAX_LEVELS}")
    dev, dt = levels[0].device, levels[0].dtype
    b = levels[0].shape[0]
    no = 4 * reg_max + nc
    d = _cabi.HeadDesc()
    d.num_levels, d.batch, d.nc, d.reg_max, d.dtype = len(levels), b, nc, reg_max, _cabi.dtype_code(dt)
    keep = []
    anchors = 0
    for i, lv in enumerate(levels):
        _cabi.require_cuda(lv, "head level")
        if lv.device != dev or lv.dtype != dt:
            raise ValueError("head levels must share device and dtype")
        if lv.dim() != 4 or lv.shape[0] != b or lv.shape[1] != no:
            raise ValueError(f"level {i}: expected (B={b}, {no}, H, W), got {tuple(lv.shape)}")
        h, w = lv.shape[2], lv.shape[3]
        if h * w and (lv.stride(3) != 1 or lv.stride(2) != w):
            lv = lv.contiguous()  # channels_last or sliced maps: anchors must be contiguous
        keep.append(lv)
        d.level_ptr[i] = lv.data_ptr()
        d.level_h[i], d.level_w[i] = h, w
        d.level_batch_stride[i] = lv.stride(0)
        d.level_channel_stride[i] = lv.stride(1)
        d.level_stride[i] = float(strides[i])
        anchors += h * w
    return d, keep, anchors


@dataclass
class NmsPlan:
    """Geometry-dependent pieces of one non_max_suppression call."""
    params: _cabi.NmsParams
    out: _cabi.NmsOut
    rows: torch.Tensor
    idx: torch.Tensor
    count: torch.Tensor
    cand: torch.Tensor
    packed: torch.Tensor  # rows and count are views of this one fp32 buffer: [B*max_det*cols rows | B counts (int32 bits)]
    scratch: torch.Tensor = None
    keep_alive: tuple = ()
    xforms: torch.Tensor = None  # (B, 8) ypb_scale_xform array when the gather rescales to the original images
    peers: object = None         # dist.PeerGather when the kernel also stores the results into every peer's buffer
    scratch_bytes: int = 0
    counters: torch.Tensor = None  # int32[B + 1] clean-on-exit row / octet counters (ypb_nms_params.clean_counters)
    count_host: torch.Tensor = None  # pinned int32[B]: the suppression kernel stores the counts here too (ypb_nms_out.count_host)


_PLAN_CACHE_MAX = 32


def _plan_cache():
    cache = getattr(_tls, "plans", None)
    if cache is None:
        from collections import OrderedDict

        cache = _tls.plans = OrderedDict()
    return cache


def make_plan(device, batch: int, anchors: int, nc: int, extra: int, conf_t: float, iou_eff: float, max_det: int,
              max_nms: int, max_wh: float, multi_label: bool, rule: int, classes=None, with_scale: bool = False,
              scale_padding: bool = True, peer_gather_group=None, nms_box=None, boxes_xyxy: bool = False,
              pad_output: bool = False, conf_per_image: torch.Tensor | None = None, rows_cap: int | None = None,
              cached: bool = False, scan_kernel: int = 0, host_counts: bool | None = None) -> NmsPlan:
    """cached=True: the plan (parameter structs, fixed-stride result buffers) is kept per (thread, stream, geometry,
    parameters) and REUSED by the next call with the same key - only for callers that copy the results out before
    returning (``split_results`` packs them into fresh tensors); everything is stream-ordered, so reuse is safe.
    host_counts (default: = cached): the plan owns a pinned host int32[B] that the suppression kernel writes the per-image
    counts into directly (mapped memory), so reading them needs a stream synchronisation but no copy; only for plans that
    persist - a pinned allocation per call would cost more than the copy it saves."""
    if host_counts is None:
        host_counts = cached
    key = None
    if cached and peer_gather_group is None and conf_per_image is None:
        ckey = None if classes is None else tuple(int(c) for c in (classes.tolist() if isinstance(classes, torch.Tensor) else classes))
        key = (device.index, torch.cuda.current_stream(device).cuda_stream, batch, anchors, nc, extra, conf_t, iou_eff, int(max_det),
               int(max_nms), float(max_wh), bool(multi_label), rule, ckey, bool(with_scale), bool(scale_padding),
               None if nms_box is None else tuple(nms_box), bool(boxes_xyxy), bool(pad_output), rows_cap, scan_kernel)
        cache = _plan_cache()
        hit = cache.get(key)
        if hit is not None:
            cache.move_to_end(key)
            hit.scratch = _scratch(device, hit.scratch_bytes)  # the pool buffer may have been regrown since
            return hit
    if key is not None and torch.is_inference_mode_enabled():
        # tensors of a CACHED plan outlive this call: create them as normal tensors, so a later call outside
        # inference_mode may still update them in place (set_transforms)
        with torch.inference_mode(False):
            plan = make_plan(device, batch, anchors, nc, extra, conf_t, iou_eff, max_det, max_nms, max_wh, multi_label, rule,
                             classes, with_scale, scale_padding, None, nms_box, boxes_xyxy, pad_output, None, rows_cap, False, scan_kernel,
                             host_counts)
        cache = _plan_cache()
        cache[key] = plan
        while len(cache) > _PLAN_CACHE_MAX:
            cache.popitem(last=False)
        return plan
    rows_cap = (anchors * nc if multi_label else anchors) if rows_cap is None else int(rows_cap)
    rows_cap = max(rows_cap, 1)
    max_nms = max(1, min(int(max_nms), rows_cap))
    max_det = max(1, min(int(max_det), max_nms))
    lib = _cabi.load()
    nbytes = lib.ypb_nms_workspace_bytes(batch, anchors, rows_cap, max_det, max_nms, rule)
    scratch = _scratch(device, nbytes)
    cols = 6 + extra
    nrow = batch * max_det * cols
    peers = None
    if peer_gather_group is not None:  # results land in slot `rank` of the node-wide symmetric buffer (dist.PeerGather)
        from .dist import PeerGather

        peers = PeerGather(nrow + batch, nrow, device, None if peer_gather_group is True else peer_gather_group)
        packed = peers.my_packed
    else:
        packed = torch.empty((nrow + batch,), dtype=torch.float32, device=device)
    rows = packed[:nrow].view(batch, max_det, cols)
    count = packed[nrow:].view(torch.int32)
    idx = torch.empty((batch, max_det), dtype=torch.int64, device=device)
    cand = torch.empty((batch,), dtype=torch.int32, device=device)
    mask = class_mask_tensor(classes, nc, device)
    p = _cabi.NmsParams()
    p.conf_thres, p.iou_thres_eff = conf_t, iou_eff
    p.nc, p.extra, p.max_det, p.max_nms = nc, extra, max_det, max_nms
    p.max_wh, p.multi_label, p.rule, p.rows_cap = float(max_wh), int(bool(multi_label)), rule, rows_cap
    p.class_mask = mask.data_ptr() if mask is not None else None
    if nms_box is not None:  # exporter NMSModel flavour: suppression on multiplier * (box / divisor)
        p.nms_box_divisor, p.nms_box_multiplier = float(nms_box[0]), float(nms_box[1])
    p.boxes_xyxy, p.pad_output = int(bool(boxes_xyxy)), int(bool(pad_output))
    p.scan_kernel = int(scan_kernel)
    if conf_per_image is not None:
        p.conf_per_image = conf_per_image.data_ptr()
    # clean-on-exit counters owned by the plan: the calls then enqueue kernels only (no memset node); zeroed here, left
    # zeroed by the suppression kernel of every completed call (run_* re-zero them if a call fails half-way)
    counters = torch.zeros((batch + 1,), dtype=torch.int32, device=device)
    p.clean_counters = counters.data_ptr()
    o = _cabi.NmsOut()
    o.rows, o.idx, o.count, o.cand_count = rows.data_ptr(), idx.data_ptr(), count.data_ptr(), cand.data_ptr()
    xforms = None
    if with_scale:  # fused construct_result rescale (detect/predict.py:120, obb/predict.py:59-60)
        xforms = torch.zeros((max(batch, 1), 8), dtype=torch.float32, device=device)
        xforms[:, 0] = 1.0
        o.scale_xforms, o.scale_padding = xforms.data_ptr(), int(bool(scale_padding))
    if peers is not None:
        peers.bind(o)
    plan = NmsPlan(p, o, rows, idx, count, cand, packed, scratch, (mask, conf_per_image, counters), xforms, peers)
    plan.counters = counters
    if host_counts and batch > 0:
        plan.count_host = torch.zeros((batch,), dtype=torch.int32).pin_memory()
        o.count_host = plan.count_host.data_ptr()  # unified addressing: the pinned allocation is mapped at the same address
    plan.scratch_bytes = nbytes
    if key is not None:
        cache = _plan_cache()
        cache[key] = plan
        while len(cache) > _PLAN_CACHE_MAX:
            cache.popitem(last=False)
    return plan


def set_transforms(plan: NmsPlan, img1_shape, orig_shapes, ratio_pads=None) -> None:
    """Load the per-image letterbox transforms of this batch into the plan's static device array (stream-ordered)."""
    from . import ops

    if plan.xforms is None:
        raise RuntimeError("the plan was built without with_scale=True")
    if len(orig_shapes) != plan.xforms.shape[0]:
        raise ValueError(f"{len(orig_shapes)} original shapes for a batch of {plan.xforms.shape[0]}")
    plan.xforms.copy_(ops.transforms_tensor(img1_shape, orig_shapes, ratio_pads, plan.xforms.device), non_blocking=True)


def fetch_counts(count: torch.Tensor) -> list:
    """The one device->host transfer of the path: per-image kept counts."""
    host = _pinned_counts(count.numel())
    host.copy_(count, non_blocking=True)
    torch.cuda.current_stream(count.device).synchronize()
    return host.tolist()


def compact_results(plan: NmsPlan, with_idx: bool, out_rows: torch.Tensor | None = None, out_idx: torch.Tensor | None = None):
    """Pack the kept rows (and anchor indices) of the plan's fixed-stride buffers back to back into fresh (or given) tensors:
    one launch, no host synchronisation."""
    b, md, cols = plan.rows.shape
    dev = plan.rows.device
    if out_rows is None:
        out_rows = torch.empty((b * md, cols), dtype=torch.float32, device=dev)
    if with_idx and out_idx is None:
        out_idx = torch.empty((b * md,), dtype=torch.int64, device=dev)
    rc = _cabi.load().ypb_compact_results(plan.rows.data_ptr(), plan.idx.data_ptr() if with_idx else None, plan.count.data_ptr(),
                                          b, md, cols, out_rows.data_ptr(), out_idx.data_ptr() if with_idx else None, None,
                                          _cabi.stream_ptr(dev))
    _cabi.check(rc, "ypb_compact_results")
    return out_rows, out_idx


def cut_results(out_rows: torch.Tensor, out_idx, counts: list, return_idxs: bool):
    """list[Tensor(n_i, cols)] (nms.py:159-161) from the packed rows: ONE split, not B slicing calls."""
    total = sum(counts)
    out = list(torch.split(out_rows[:total], counts)) if counts else []
    if return_idxs:
        return out, (list(torch.split(out_idx[:total], counts)) if counts else [])
    return out


def plan_counts(plan: NmsPlan) -> list:
    """Per-image kept counts of the plan's last call on the host: one stream synchronisation; no copy when the kernel wrote
    them into the plan's mapped host buffer."""
    if plan.count_host is None:
        return fetch_counts(plan.count)
    torch.cuda.current_stream(plan.count.device).synchronize()
    return plan.count_host.tolist()


def split_results(plan: NmsPlan, return_idxs: bool):
    if plan.rows.shape[0] == 0:
        return ([], []) if return_idxs else []
    out_rows, out_idx = compact_results(plan, return_idxs)
    return cut_results(out_rows, out_idx, plan_counts(plan), return_idxs)


def _check_plan(rc: int, what: str, plan: NmsPlan) -> None:
    if rc != 0 and plan.counters is not None:
        plan.counters.zero_()  # a call that failed half-way may have left rows counted: restore the clean-on-exit invariant
    _cabi.check(rc, what)


def run_from_dense(pred: torch.Tensor, plan: NmsPlan, anchor_subset: torch.Tensor | None = None) -> None:
    """anchor_subset: optional (B, n) int64 device tensor - only these anchors of ``pred`` are candidates (read in place)."""
    d = _cabi.DenseDesc()
    d.ptr, d.dtype = pred.data_ptr(), _cabi.dtype_code(pred.dtype)
    d.batch, d.channels, d.anchors = pred.shape
    d.stride_b, d.stride_c, d.stride_a = pred.stride()
    if anchor_subset is not None:
        if anchor_subset.dtype != torch.int64 or anchor_subset.dim() != 2 or anchor_subset.shape[0] != pred.shape[0] \
                or not anchor_subset.is_contiguous() or anchor_subset.device != pred.device:
            raise ValueError("anchor_subset must be a contiguous (B, n) int64 tensor on the predictions' device")
        d.anchor_subset, d.subset_len = anchor_subset.data_ptr(), anchor_subset.shape[1]
    lib = _cabi.load()
    rc = lib.ypb_nms_from_dense(C.byref(d), C.byref(plan.params), C.byref(plan.out), plan.scratch.data_ptr(),
                                plan.scratch.numel(), _cabi.stream_ptr(pred.device))
    _check_plan(rc, "ypb_nms_from_dense", plan)


def geometry_desc(level_hw, strides, batch: int, dtype):
    """ypb_head_desc carrying only the level geometry (no tensors): for calls that need anchors -> grid cells."""
    d = _cabi.HeadDesc()
    d.num_levels, d.batch, d.nc, d.reg_max, d.dtype = len(level_hw), batch, 1, 16, _cabi.dtype_code(dtype)
    for i, ((h, w), s) in enumerate(zip(level_hw, strides)):
        d.level_h[i], d.level_w[i], d.level_stride[i] = int(h), int(w), float(s)
    return d


def riders_desc(t: torch.Tensor, batch: int, anchors: int, dtype, kind: int, kpt_ndim: int = 0):
    """ypb_riders_desc for a (B, E, A) tensor of per-anchor channels (mask coefficients / raw keypoints)."""
    _cabi.require_cuda(t, "riders")
    if t.dim() != 3 or t.shape[0] != batch or t.shape[2] != anchors:
        raise ValueError(f"riders must be (B={batch}, E, A={anchors}), got {tuple(t.shape)}")
    if t.dtype != dtype:
        t = t.to(dtype)
    if t.stride(2) != 1:
        t = t.contiguous()
    r = _cabi.RidersDesc()
    r.ptr, r.channels, r.kind, r.kpt_ndim = t.data_ptr(), t.shape[1], kind, kpt_ndim
    r.stride_b, r.stride_c = t.stride(0), t.stride(1)
    return r, t


def run_from_head_riders(desc, riders, plan: NmsPlan, device) -> None:
    lib = _cabi.load()
    rc = lib.ypb_nms_from_head_riders(C.byref(desc), C.byref(riders), desc.dtype, C.byref(plan.params), C.byref(plan.out),
                                      plan.scratch.data_ptr(), plan.scratch.numel(), _cabi.stream_ptr(device))
    _check_plan(rc, "ypb_nms_from_head_riders", plan)


def run_from_head(desc, angle, angle_is_logit: bool, plan: NmsPlan, device) -> None:
    lib = _cabi.load()
    rc = lib.ypb_nms_from_head(C.byref(desc), angle.data_ptr() if angle is not None else None, int(angle_is_logit),
                               desc.dtype, C.byref(plan.params), C.byref(plan.out), plan.scratch.data_ptr(),
                               plan.scratch.numel(), _cabi.stream_ptr(device))
    _check_plan(rc, "ypb_nms_from_head", plan)
