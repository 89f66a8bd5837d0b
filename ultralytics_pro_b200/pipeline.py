"""Fixed-geometry post-processor: descriptors, scratch and result buffers built once, optional CUDA-graph replay.

``postprocess_from_head`` rebuilds its plan on every call (fine for a predictor loop).  A serving loop that sees the
same head geometry every step uses this class instead: one C-ABI call (or one graph launch) per batch, no host
allocation, no host synchronisation until ``results()``.  One instance serves one stream at a time (its result
buffers and scratch are reused from call to call).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi, engine
from .nms import _greedy_threshold


class HeadPostProcessor:
    def __init__(self, nc: int, strides, conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None,
                 agnostic: bool = False, multi_label: bool = False, max_det: int = 300, max_nms: int = 30000,
                 max_wh: int = 7680, reg_max: int = 16, rotated: bool = False, scale_to_original: bool = False,
                 peer_gather_group=None, use_graph: bool = False, scan_kernel: str = "auto"):
        assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
        assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
        self.nc, self.strides, self.reg_max = nc, tuple(float(s) for s in strides), reg_max
        self.conf_thres, self.iou_thres = float(conf_thres), float(iou_thres)
        self.classes, self.agnostic, self.multi_label = classes, agnostic, bool(multi_label) and nc > 1
        self.max_det, self.max_nms, self.max_wh, self.rotated = max_det, max_nms, max_wh, rotated
        # fold construct_result's scale_boxes / regularize_rboxes (detect/predict.py:120, obb/predict.py:59-60) into the
        # gather; the per-image transforms are loaded with set_image_shapes() before enqueue() / graph replay
        self.scale_to_original = bool(scale_to_original)
        # multi-GPU: True / a process group = one-sided gather of every rank's results over NVLink peer memory
        # (dist.PeerGather); plans are then created collectively, in the same order on every rank
        self.peer_gather_group = peer_gather_group
        # use_graph: __call__ replays ONE CUDA graph per distinct set of input tensors (kernels + result packing; the counts
        # land in mapped pinned memory straight from the suppression kernel): a serving loop with static input buffers pays
        # one graph launch and one stream synchronisation per batch.  The returned views are then valid until the next call on the same inputs.
        self.use_graph = bool(use_graph)
        # class-scan kernel of the fused path: "ldg" (one-wave grid, co-resident with other streams' kernels: multi-stream
        # pipelines), "tma" (persistent TMA-fed ring: fastest single kernel, latency mode / 16-bit heads), "auto" = ldg
        self.scan_kernel = {"auto": _cabi.SCAN_AUTO, "ldg": _cabi.SCAN_LDG, "tma": _cabi.SCAN_TMA}[scan_kernel]
        self._graphs = {}
        self._plans = {}
        self.last = None

    def _plan_for(self, levels, anchors):
        lv0 = levels[0]
        key = (lv0.device.index, lv0.dtype, lv0.shape[0], anchors)
        plan = self._plans.get(key)
        if plan is None:
            conf_t = _cabi.round_to_dtype(self.conf_thres, lv0.dtype)
            if self.rotated:
                rule, iou_eff = _cabi.RULE_FAST_PROBIOU, _cabi.f32_round(self.iou_thres)
            else:
                rule, iou_eff = _cabi.RULE_GREEDY, _greedy_threshold(self.iou_thres)
            with torch.inference_mode(False):  # the plan's tensors outlive this call (set_image_shapes updates them in place)
                plan = engine.make_plan(lv0.device, lv0.shape[0], anchors, self.nc, 1 if self.rotated else 0, conf_t,
                                        iou_eff, self.max_det, self.max_nms, 0.0 if self.agnostic else float(self.max_wh),
                                        self.multi_label, rule, self.classes, with_scale=self.scale_to_original,
                                        peer_gather_group=self.peer_gather_group, scan_kernel=self.scan_kernel,
                                        host_counts=True)
                # a private scratch buffer: the plan outlives the call, the thread-local pool buffer may be regrown
                nbytes = _cabi.load().ypb_nms_workspace_bytes(lv0.shape[0], anchors, plan.params.rows_cap,
                                                              plan.params.max_det, plan.params.max_nms, plan.params.rule)
                plan.scratch = torch.empty(nbytes, dtype=torch.uint8, device=lv0.device)
            self._plans[key] = plan
        return plan

    def set_image_shapes(self, levels, img_shape, orig_shapes, ratio_pads=None):
        """Stream-ordered update of the static per-image transform array of the plan for this geometry (graph-safe:
        the captured kernels read the array at replay time)."""
        _, keep, anchors = engine.head_desc(levels, self.strides, self.nc, self.reg_max)
        engine.set_transforms(self._plan_for(keep, anchors), img_shape, orig_shapes, ratio_pads)

    def enqueue(self, levels, angle_logits=None, stage: int = 0):
        """Launch decode+NMS for one batch on the current stream; returns the device-resident plan (no sync)."""
        desc, keep, anchors = engine.head_desc(levels, self.strides, self.nc, self.reg_max)
        plan = self._plan_for(keep, anchors)
        ang = None
        if self.rotated:
            if angle_logits is None:
                raise ValueError("rotated post-processor needs angle_logits")
            ang = angle_logits.reshape(keep[0].shape[0], anchors)
            if not ang.is_contiguous() or ang.dtype != keep[0].dtype:
                ang = ang.to(keep[0].dtype).contiguous()
        lib = _cabi.load()
        dev = keep[0].device
        rc = lib.ypb_nms_from_head_stage(C.byref(desc), ang.data_ptr() if ang is not None else None, 1, desc.dtype,
                                         C.byref(plan.params), C.byref(plan.out), plan.scratch.data_ptr(),
                                         plan.scratch.numel(), _cabi.stream_ptr(dev), stage)
        engine._check_plan(rc, "ypb_nms_from_head_stage", plan)
        self.last = plan
        return plan

    def capture(self, levels, angle_logits=None, after=None, beside=None) -> "torch.cuda.CUDAGraph":
        """Capture one batch's launches into a CUDA graph bound to these input tensors (static addresses).
        `after()` (optional) is captured behind the kernels, e.g. the NCCL gather of the result buffer.
        `beside()` (optional) is captured on a FORKED branch that runs concurrently with the kernels and joins at the end of
        the graph - e.g. the consumer of the PREVIOUS batch's one-sided gather (``wait_gather(-1)`` + its reads), which then
        costs the step no latency at all."""
        if beside is not None:
            beside()  # eagerly too: the warm-up launch below must find the sequence the replays will continue
        self.enqueue(levels, angle_logits)  # warm: attributes set, plan built, scratch allocated
        if after is not None:
            after()
        dev = levels[0].device
        cur = torch.cuda.current_stream(dev)
        cur.synchronize()
        graph = torch.cuda.CUDAGraph()
        # capture needs a non-default stream; keep the caller's stream when it already is one
        cap = cur if cur != torch.cuda.default_stream(dev) else torch.cuda.Stream(dev)
        side = torch.cuda.Stream(dev) if beside is not None else None
        with torch.cuda.graph(graph, stream=cap):
            if side is not None:
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    beside()
            self.enqueue(levels, angle_logits)
            if after is not None:
                after()
            if side is not None:
                torch.cuda.current_stream(dev).wait_stream(side)
        return graph

    def wait_gather(self, lag: int = 0):
        """Enqueue the consumer-side wait of the one-sided gather (graph-capturable): lag=0 for the last enqueued batch,
        lag=k for the batch k enqueues back on this processor (pipelined: never stalls on a slower rank), lag=-1 for the batch
        after the one handed out last (in order - for a consumer running beside the next batch's kernels, ``capture(beside=)``)."""
        if self.last is None or self.last.peers is None:
            raise RuntimeError("no peer gather attached to the last plan")
        self.last.peers.wait(lag)

    def gathered(self):
        """(world*B, max_det, cols) rows and (world*B,) counts of all ranks for the last waited-for batch."""
        pl = self.last
        return pl.peers.gathered(pl.rows.shape[0], pl.rows.shape[1], pl.rows.shape[2])

    def __call__(self, levels, angle_logits=None, return_idxs: bool = False):
        if self.use_graph:
            return self._call_graphed(levels, angle_logits, return_idxs)
        return engine.split_results(self.enqueue(levels, angle_logits), return_idxs)

    def _call_graphed(self, levels, angle_logits, return_idxs: bool):
        dev = levels[0].device
        key = (tuple((lv.data_ptr(), tuple(lv.shape), lv.stride()) for lv in levels), levels[0].dtype,
               angle_logits.data_ptr() if angle_logits is not None else 0)
        ent = self._graphs.get(key)
        if ent is None:
            plan = self.enqueue(levels, angle_logits)  # warm-up: plan, scratch and result buffers exist after this
            single = plan.rows.shape[0] == 1  # one image: its kept rows already lie back to back - nothing to pack
            out_rows = out_idx = None
            if single:
                out_rows, out_idx = plan.rows[0], plan.idx[0]
            else:
                with torch.inference_mode(False):
                    out_rows, out_idx = engine.compact_results(plan, True)
            cur = torch.cuda.current_stream(dev)
            cur.synchronize()
            graph = torch.cuda.CUDAGraph()
            cap = cur if cur != torch.cuda.default_stream(dev) else torch.cuda.Stream(dev)
            with torch.cuda.graph(graph, stream=cap):
                self.enqueue(levels, angle_logits)
                if not single:
                    engine.compact_results(plan, True, out_rows, out_idx)
                # no count copy: the suppression kernel writes the counts into the plan's mapped host buffer itself
            if len(self._graphs) >= 8:
                self._graphs.pop(next(iter(self._graphs)))
            ent = self._graphs[key] = (graph, plan, out_rows, out_idx, plan.count_host, list(levels), angle_logits)
        graph, plan, out_rows, out_idx, host = ent[:5]
        graph.replay()
        torch.cuda.current_stream(dev).synchronize()
        self.last = plan
        return engine.cut_results(out_rows, out_idx, host.tolist(), return_idxs)

    def results(self, return_idxs: bool = False):
        if self.last is None:
            raise RuntimeError("nothing enqueued yet")
        return engine.split_results(self.last, return_idxs)
