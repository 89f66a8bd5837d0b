#!/usr/bin/env python
"""Benchmark of the detection post-processing hot path (BASELINE.json: post-process imgs/s, decode+NMS, B=64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype f32|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the fused decode+filter+sort+suppress path over one batch of synthetic yolov8x-shaped head
outputs (configs[1] of BASELINE.json) per GPU.  Prints ONE JSON line on rank 0.  See DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "post-process imgs/s (decode+NMS) @B=64"
UNIT = "imgs/s"
WORKLOAD = "c2_v8x_640_b64"
WORKLOAD_DESC = ("c2: yolov8x detect head outputs (3 levels 80/40/20, 144 ch, 8400 anchors, 80 classes, reg_max 16), "
                 "batch 64 per GPU at 640x640, predict mode conf=0.25 iou=0.7 max_det=300, clustered-object synthetic logits")
FALLBACK_HBM_GBS = 6650.0


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def _traffic(dtype):
    """dram bytes per launch of the roofline kernel from the committed ncu capture (profiles/roofline_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            return json.load(fh).get(dtype)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20}
    NOTE = {"sw_power_cap": 0x4, "hw_power_brake": 0x80}

    def __init__(self, index: int, period: float = 0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            uuid = None
            try:
                uuid = torch.cuda.get_device_properties(index).uuid
            except Exception:
                pass
            self.nv = pynvml
            self.h = None
            if uuid is not None:
                try:
                    self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
                except Exception:
                    self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self):
        if not self.ok:
            return
        try:
            self.samples.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(
                self.nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(
                self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            for name, bit in {**self.BAD, **self.NOTE}.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while not self._stop_evt.is_set():
            self.sample()
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        self.sample()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


# ----------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port (torch CPU restatement of the reference's own operators) on host cores
# ----------------------------------------------------------------------------------------------------------------
def _cpu_step_fn(cfg, images: int, seed: int = 99):
    from oracle.postproc_oracle import decode_oracle, nms_oracle
    from ultralytics_pro_b200.synth import make_head_batch

    levels, _ = make_head_batch(cfg, batch=images, seed=seed)

    def step():
        y = decode_oracle(levels, cfg.strides, cfg.nc, cfg.reg_max)
        out, _ = nms_oracle(y, cfg.conf, cfg.iou, nc=cfg.nc, multi_label=cfg.multi_label, max_det=cfg.max_det)
        return out

    return step


def cpu_baseline(cfg, budget_s: float = 15.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = _cpu_step_fn(cfg, cfg.batch)
    step()  # warm-up (imports torchvision)
    reps, t0 = 0, time.perf_counter()
    while True:
        step()
        reps += 1
        el = time.perf_counter() - t0
        if el > budget_s or reps >= 50:
            break
    all_cores = cfg.batch * reps / el
    torch.set_num_threads(1)
    t1 = time.perf_counter()
    step()
    one = cfg.batch / (time.perf_counter() - t1)
    torch.set_num_threads(cores)
    return {"value": all_cores, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{reps} passes over one {cfg.batch}-image batch of the same workload (decode_oracle + nms_oracle, "
                      f"torchvision.ops.nms branch), torch threads={cores}",
            "single_thread_value": one}


def run_reference(args):
    rank, world, _ = _dist_env()
    if rank != 0:
        return 0
    from ultralytics_pro_b200.synth import CONFIGS

    cfg = CONFIGS[WORKLOAD]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    probe = _cpu_step_fn(cfg, 8)
    probe()
    t0 = time.perf_counter()
    probe()
    t_img = (time.perf_counter() - t0) / 8
    total_steps = args.steps + args.warmup
    sample = int(max(1, min(cfg.batch, 120.0 / max(t_img * total_steps, 1e-9))))
    step = _cpu_step_fn(cfg, sample)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    el = time.perf_counter() - t0
    value = sample * args.steps / el
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC, "batch_per_step": sample, "device": "host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} images of the workload per step (bounded so the run ends in minutes); "
                                   f"oracle port = the reference's torch CPU operators incl. torchvision.ops.nms"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def _post_for(cfg, gather_mode, scan_kernel="auto"):
    from ultralytics_pro_b200.pipeline import HeadPostProcessor

    return HeadPostProcessor(cfg.nc, cfg.strides, cfg.conf, cfg.iou, multi_label=cfg.multi_label, agnostic=cfg.agnostic,
                             rotated=cfg.rotated, max_det=cfg.max_det, max_nms=cfg.max_nms,
                             peer_gather_group=True if gather_mode == "peer" else None, scan_kernel=scan_kernel)


def _build_lanes(dev, cfg, sets, nlanes, world, gather_mode, lag, use_graph, lane_groups=None, scan_kernel="auto", consumer="beside"):
    """`nlanes` independent pipelines (own stream, plan, scratch, result buffers, CUDA graphs); lane i owns input sets i and
    i + nlanes.  One step of a lane = [class scan, survivor decode, sort+suppress (+ peer push)], then - all
    inside the lane's CUDA graph - the consumer side: multi-GPU: wait for the gathered batch (`lag` batches back), copy the
    gathered entry out of the ring (the consumer; the entry is released by the next wait); always: the per-image counts to
    pinned HOST memory ("detections resident in HBM + counts on host")."""
    from ultralytics_pro_b200 import dist as ypb_dist

    main = torch.cuda.current_stream(dev)
    lanes = []
    for ln in range(nlanes):
        st = torch.cuda.Stream(dev)
        st.wait_stream(main)
        with torch.cuda.stream(st):
            pp = _post_for(cfg, gather_mode, scan_kernel)
            my_sets = [sets[(ln + nlanes * j) % len(sets)] for j in range(2)]
            pl = pp.enqueue(*my_sets[0])  # builds the plan / result buffers (collective when the peer gather is on)
            # the per-image counts reach pinned HOST memory inside every step: written there by the suppression kernel itself
            # (ypb_nms_out.count_host, mapped memory) - no copy node; a plan without that buffer gets an explicit copy
            host_counts = pl.count_host if pl.count_host is not None else torch.empty((pl.count.numel(),), dtype=torch.int32).pin_memory()
            gb = None
            if gather_mode in ("nccl", "nccl-eager"):
                gb = torch.empty((world, pl.packed.numel()), dtype=torch.float32, device=dev)
            elif gather_mode == "peer":
                gb = torch.empty((1, world, pl.peers.slot), dtype=torch.float32, device=dev)  # the consumer's copy of the entry
            grp = lane_groups[ln] if lane_groups else None

            # consumer = "beside": the gathered batch q-1 is consumed on a forked branch of step q's graph; "tail": behind step q-1's
            # suppression kernel (round 1's placement)
            beside_mode = gather_mode == "peer" and use_graph and lag == 1 and consumer == "beside"

            def consume(pp=pp, pl=pl, gb=gb, want=(-1 if beside_mode else lag)):
                pp.wait_gather(want)
                if not os.environ.get("YPB_BENCH_NO_CONSUME"):
                    # the consumer: one device-side gather of the returned ring entry (measured faster than the single-CTA
                    # ypb_peer_wait_copy for this 0.9 MB entry: 46.0 vs 50.9 us per step at N=2)
                    pl.peers.copy_entry(gb)

            def tail(pp=pp, pl=pl, gb=gb, host_counts=host_counts, grp=grp):
                if gather_mode == "peer" and not beside_mode:
                    consume()
                elif gather_mode == "nccl":
                    ypb_dist.gather_packed(pl.packed, gb, group=grp)
                if pl.count_host is None:
                    host_counts.copy_(pl.count, non_blocking=True)

            graphs = None
            if use_graph:
                # lag 1 + graphs: the consumer of batch q-1 (in-order wait + copy out of the ring) is a FORKED branch of step q's
                # graph - it runs beside the class scan of step q instead of behind its suppression kernel
                graphs = [pp.capture(lv, ang, after=tail, beside=consume if beside_mode else None) for lv, ang in my_sets]
        st.synchronize()
        lanes.append({"stream": st, "post": pp, "sets": my_sets, "graphs": graphs, "plan": pl, "gather": gb, "tail": tail, "beside": beside_mode, "primed": True,
                      "group": grp, "host_counts": host_counts})
    return lanes


def _lane_step(lanes, i, gather_mode):
    from ultralytics_pro_b200 import dist as ypb_dist

    ln = lanes[i % len(lanes)]
    j = (i // len(lanes)) % 2
    with torch.cuda.stream(ln["stream"]):
        if ln["graphs"] is not None:
            ln["graphs"][j].replay()
        else:
            ln["post"].enqueue(*ln["sets"][j])
            ln["tail"]()
        if gather_mode == "nccl-eager":
            ypb_dist.gather_packed(ln["plan"].packed, ln["gather"], group=ln["group"])


class _DeviceBarrier:
    """Optional (diagnostic) GPU-side barrier over NVLink peer memory (torch symmetric memory), enqueued on the current stream.
    dist.barrier() lines the HOSTS up to within tens of microseconds; this lines the GPUs themselves up before the first event
    is recorded.  Collective construction."""

    def __init__(self, dev, world):
        self.handle = None
        # OFF unless YPB_BENCH_DEVICE_BARRIER=1.  Measured (profiles/r02_n4_k20.txt, r02_n2_k20.txt): no gain at N=2 (48.1 vs 48.0 us per
        # step at K=20), a LOSS at N=4 (49.4 vs 47.2) and N=8 (59.9) - GPUs that start in the same microsecond issue their peer stores
        # and flag writes in the same microseconds for the whole short run; the hosts' natural skew spreads them.
        if world <= 1 or not os.environ.get("YPB_BENCH_DEVICE_BARRIER"):
            return
        ok = 1
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm

            self.buf = symm.empty(64, dtype=torch.float32, device=dev)
            try:
                self.handle = symm.rendezvous(self.buf, dist.group.WORLD)
            except Exception:  # noqa: BLE001
                symm.enable_symm_mem_for_group(dist.group.WORLD.group_name)
                self.handle = symm.rendezvous(self.buf, dist.group.WORLD)
            self.handle.barrier(channel=0)
            torch.cuda.synchronize(dev)
        except Exception as exc:  # noqa: BLE001
            sys.stderr.write(f"[bench] device-side barrier unavailable ({type(exc).__name__}: {exc})\n")
            ok = 0
        import torch.distributed as dist

        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            self.handle = None

    def __call__(self):
        if self.handle is not None:
            self.handle.barrier(channel=0)


def _time_lanes(dev, lanes, K, W, world, gather_mode, lag, sampler=None, device_barrier=None):
    """W warm-up steps, then EXACTLY K timed steps between a barrier + synchronize on both sides; CUDA events on the stream the
    lanes fork from / join into; max over ranks.  Returns total ms."""
    import torch.distributed as dist

    main = torch.cuda.current_stream(dev)

    def fork():
        for ln in lanes:
            ln["stream"].wait_stream(main)

    def join():
        for ln in lanes:
            main.wait_stream(ln["stream"])

    def drain(final=True):
        if gather_mode == "peer" and lag > 0:  # every result of every rank has landed before the clock stops
            if not final and lanes[0].get("beside"):
                return  # in-order consumers: the warm-up leaves the steady state (one batch in flight per lane) in place
            for ln in lanes:
                with torch.cuda.stream(ln["stream"]):
                    ln["post"].wait_gather(0)
                ln["primed"] = False  # every batch has been handed out: an in-order consumer would now wait for its OWN step's launch

    fork()
    for ln in lanes:
        if ln.get("beside") and not ln["primed"]:
            # restore the steady state of the lag-1 pipeline after a drain: one launch in flight that nobody has consumed yet
            with torch.cuda.stream(ln["stream"]):
                ln["post"].enqueue(*ln["sets"][0])
            ln["primed"] = True
    for i in range(W):
        _lane_step(lanes, i, gather_mode)
    drain(final=False)
    join()
    torch.cuda.synchronize(dev)
    if sampler is not None:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    if device_barrier is not None:
        device_barrier()  # the GPUs themselves leave this point together (see _DeviceBarrier)
    ev0.record()
    fork()
    for i in range(K):
        _lane_step(lanes, i, gather_mode)
    drain()
    join()
    ev1.record()
    torch.cuda.synchronize(dev)
    if sampler is not None:
        sampler.stop()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def _stage_times(dev, post, sets, K):
    """Per-kernel durations with CUDA events on the launching stream: each prefix of the step (scan | scan+decode |
    scan+decode+suppress) is captured as a CUDA graph of R back-to-back repetitions over rotating input sets (> L2) and
    replayed between two events - no host launch gaps inside the timed region; durations = differences of the prefixes."""
    R = len(sets)
    tstream = torch.cuda.Stream(dev)
    prefix_ms = []
    with torch.cuda.stream(tstream):
        for mask in (1, 3, 7):
            post.enqueue(*sets[0], stage=mask)
            tstream.synchronize()
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=tstream):
                for r in range(R):
                    post.enqueue(*sets[r % R], stage=mask)
            for _ in range(3):
                gph.replay()
            tstream.synchronize()
            KI = max(3, min(K, 200) // R)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(tstream)
            for _ in range(KI):
                gph.replay()
            b.record(tstream)
            tstream.synchronize()
            prefix_ms.append(a.elapsed_time(b) / (KI * R))
            del gph
        # the class-scan KERNEL alone (stage 1 | 8: no counter memset in front of every launch; the counters are cleared once
        # per graph replay, so the rows of the R launches accumulate in the key buffer - R x ~600 rows of 8400 slots per image)
        plan = post.last
        if plan is not None and plan.counters is not None:
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph, stream=tstream):
                plan.counters.zero_()
                for r in range(R):
                    post.enqueue(*sets[r % R], stage=9)
            for _ in range(3):
                gph.replay()
            tstream.synchronize()
            KI = max(3, min(K, 200) // R)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(tstream)
            for _ in range(KI):
                gph.replay()
            b.record(tstream)
            tstream.synchronize()
            prefix_ms.append(a.elapsed_time(b) / (KI * R))
            plan.counters.zero_()
            tstream.synchronize()
            del gph
    return prefix_ms


def _dense_decode_time(dev, cfg, sets, esize, peak, reps):
    from ultralytics_pro_b200.head import decode_head

    B = sets[0][0][0].shape[0]

    def f(i):
        lv, ang = sets[i % len(sets)]
        if cfg.rotated:
            return decode_head(lv, cfg.strides, cfg.nc, angle=ang, angle_is_logit=True, append_angle=True)
        return decode_head(lv, cfg.strides, cfg.nc)

    for i in range(3):
        y = f(i)
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    d0.record()
    for i in range(reps):
        y = f(i)
    d1.record()
    torch.cuda.synchronize(dev)
    del y
    t = d0.elapsed_time(d1) / reps
    ne = 1 if cfg.rotated else 0
    nbytes = B * cfg.anchors * ((cfg.no + ne) + (4 + cfg.nc + ne)) * esize
    return {"launch_ms": t, "achieved": nbytes / (t * 1e-3) / 1e9, "unit": "GB/s", "frac": nbytes / (t * 1e-3) / 1e9 / peak,
            "algorithmic_bytes_per_launch": nbytes}


def _e2e(dev, cfg, sets, world, KE):
    """The metric end to end through the public API with HOST buffers: every step copies that step's head tensors from
    pinned host memory, runs postprocess_from_head and reads rows + counts back to the host.  Double-buffered: the H2D copy
    of step k+1 (copy stream) runs under the kernels and the D2H of step k; the host waits for step k-1's results while
    step k is in flight - every step's inputs still cross PCIe inside the timed region."""
    import torch.distributed as dist

    from ultralytics_pro_b200.head import postprocess_from_head

    lv0 = sets[0][0]
    B = lv0[0].shape[0]
    host_sets = [[lv.cpu().pin_memory() for lv in s[0]] for s in sets[:2]]
    dev_in = [[torch.empty_like(lv) for lv in lv0] for _ in range(2)]
    copy_stream, comp = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    buf_free = [torch.cuda.Event() for _ in range(2)]
    res_done = [torch.cuda.Event() for _ in range(2)]
    md = min(cfg.max_det, cfg.anchors)
    host_rows = [torch.empty((B, md, 6), dtype=torch.float32).pin_memory() for _ in range(2)]
    host_cnt = [torch.empty((B,), dtype=torch.int32).pin_memory() for _ in range(2)]
    h2d = sum(lv.numel() * lv.element_size() for lv in lv0)
    d2h = host_rows[0].numel() * 4 + B * 4
    for e in buf_free:
        e.record(comp)

    def issue(i):
        k = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(buf_free[k])
            for dst, src in zip(dev_in[k], host_sets[k]):
                dst.copy_(src, non_blocking=True)
            h2d_done[k].record(copy_stream)
        with torch.cuda.stream(comp):
            comp.wait_event(h2d_done[k])
            p = postprocess_from_head(dev_in[k], cfg.strides, cfg.nc, cfg.conf, cfg.iou, max_det=cfg.max_det, sync=False)
            buf_free[k].record(comp)
            host_rows[k].copy_(p.rows, non_blocking=True)
            host_cnt[k].copy_(p.count, non_blocking=True)
            res_done[k].record(comp)

    def run(n):
        total = 0
        for i in range(n):
            issue(i)
            if i >= 1:
                res_done[(i - 1) % 2].synchronize()  # results of the previous step are on the host now
                total += int(host_cnt[(i - 1) % 2].sum())
        res_done[(n - 1) % 2].synchronize()
        return total + int(host_cnt[(n - 1) % 2].sum())

    run(4)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    kept = run(KE)
    torch.cuda.synchronize(dev)
    el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    sec = float(el.item())
    return {"value": world * B * KE / sec, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": KE,
            "h2d_gbs_per_gpu": h2d * KE / sec / 1e9, "kept_total": kept,
            "note": "postprocess_from_head on pinned HOST head tensors, every step: H2D of the head (copy stream, double-buffered, "
                    "overlapping the previous step's kernels), decode+NMS, D2H of rows and counts, host waits for each step's result; "
                    "wall clock, max over ranks"}


def _measure_config(dev, name, dtype, K, lanes_n=3):
    """imgs/s of the fused path on another BASELINE.json config (device-resident inputs, 2 rotating sets, CUDA graphs)."""
    from ultralytics_pro_b200.synth import CONFIGS, make_head_batch

    cfg = CONFIGS[name]
    sets = [make_head_batch(cfg, batch=cfg.batch, seed=2000 + s, device=dev, dtype=dtype) for s in range(2 * lanes_n)]
    lanes = _build_lanes(dev, cfg, sets, lanes_n, 1, "none", 0, True)
    ms = _time_lanes(dev, lanes, K, 5, 1, "none", 0)
    pl = lanes[0]["plan"]
    pre = _stage_times(dev, lanes[0]["post"], sets, K)
    esize = 4 if dtype == torch.float32 else 2
    in_bytes = cfg.batch * cfg.anchors * (cfg.no + (1 if cfg.rotated else 0)) * esize
    rec = {"imgs_per_s": cfg.batch * K / (ms * 1e-3), "ms_per_step": ms / K, "batch": cfg.batch, "lanes": lanes_n,
           "single_stream_ms": {"scan_classes": pre[0], "decode_tiles": pre[1] - pre[0], "sort_suppress": pre[2] - pre[1], "step": pre[2]},
           "cand_per_img": float(pl.cand.float().mean()), "kept_per_img": float(pl.count.float().mean()),
           "head_bytes_per_step": in_bytes, "head_gbs": in_bytes / (ms / K * 1e-3) / 1e9,
           "l2": f"{len(sets)} rotating input sets of {in_bytes / 1e6:.0f} MB"}
    del lanes, sets
    torch.cuda.empty_cache()
    return rec


def run_ours(args):
    import torch.distributed as dist

    from ultralytics_pro_b200 import dist as ypb_dist
    from ultralytics_pro_b200.head import postprocess_from_head
    from ultralytics_pro_b200.pipeline import HeadPostProcessor
    from ultralytics_pro_b200.synth import CONFIGS, make_head_batch

    rank, world, local = _dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: the path has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import datetime

        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    cfg = CONFIGS[WORKLOAD]
    dtype = {"f32": torch.float32, "bf16": torch.bfloat16, "f16": torch.float16}[args.dtype]
    B, K, W = cfg.batch, args.steps, max(args.warmup, 3)
    esize = 4 if dtype == torch.float32 else 2
    in_bytes_img = cfg.no * cfg.anchors * esize  # algorithmic bytes per image of the fused path

    # ---- synthetic inputs, resident in HBM before any timing; NSETS x 310 MB > 126 MB L2 ---------------------------
    # LANES independent pipelines (own stream, scratch, result buffers, CUDA graphs) take the steps round-robin, so the
    # HBM-bound class scan of one batch overlaps the latency-bound survivor decode / suppression of the previous one.
    LANES = max(1, args.lanes)
    NSETS = 2 * LANES
    gather_mode = args.gather if world > 1 else "none"
    if gather_mode == "peer":
        # collective probe: if the symmetric-memory mapping is unavailable on this box on ANY rank, all ranks fall back together
        ok = 1
        try:
            ypb_dist.PeerGather(1024, 1000, dev)
        except Exception as exc:  # noqa: BLE001
            sys.stderr.write(f"[bench] rank {rank}: peer-memory gather unavailable ({type(exc).__name__}: {exc}); falling back to NCCL\n")
            ok = 0
        flag = torch.tensor([ok], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            gather_mode = "nccl-eager"
    # one NCCL communicator PER LANE: collectives of different lanes are then unordered with respect to each other, so each
    # lane's all_gather can be captured inside that lane's CUDA graph (on ONE communicator, replays from several streams have
    # no defined cross-rank order and can deadlock)
    lane_groups = ([dist.new_group(backend="nccl") for _ in range(LANES)] if gather_mode in ("nccl", "nccl-eager") else None)
    sets = [make_head_batch(cfg, batch=B, seed=1000 + s, device=dev, dtype=dtype, first_image=rank * B) for s in range(NSETS)]
    use_graph = not args.no_graph
    consumer = os.environ.get("YPB_BENCH_CONSUMER", "auto")  # auto (default: a 40-step trial of each placement outside the timed region) | beside | tail
    lanes = _build_lanes(dev, cfg, sets, LANES, world, gather_mode, args.gather_lag, use_graph, lane_groups,
                         consumer="tail" if consumer == "tail" else "beside")
    consumer_pick = None
    if consumer == "auto" and lanes[0]["beside"]:
        # Where the consumer of the gathered batch sits (forked beside the next step / behind its own step) is a placement
        # choice whose effect depends on the number of ranks: try both for a few untimed steps OUTSIDE the timed region and keep the
        # faster one (all ranks decide on the same max-over-ranks numbers).  Measured at N=2 and N=4 the forked form wins.
        t_b = _time_lanes(dev, lanes, 40, 5, world, gather_mode, args.gather_lag)
        lanes_t = _build_lanes(dev, cfg, sets, LANES, world, gather_mode, args.gather_lag, use_graph, lane_groups, consumer="tail")
        t_t = _time_lanes(dev, lanes_t, 40, 5, world, gather_mode, args.gather_lag)
        consumer_pick = {"beside_us_per_step": t_b / 40 * 1e3, "tail_us_per_step": t_t / 40 * 1e3}
        if t_t < t_b:
            lanes, lanes_t = lanes_t, lanes
        consumer_pick["picked"] = "beside" if lanes[0]["beside"] else "tail"
        del lanes_t
    post, plan = lanes[0]["post"], lanes[0]["plan"]

    # everything with a variable host cost (NVML init of the clock sampler, event creation) happens BEFORE the barrier inside
    # _time_lanes, so that the ranks enter the timed region together
    sampler = ClockSampler(local)
    dev_barrier = _DeviceBarrier(dev, world)
    total_ms = _time_lanes(dev, lanes, K, W, world, gather_mode, args.gather_lag, sampler, dev_barrier)
    sys.stderr.write(f"[bench] rank {rank}: {total_ms / K * 1e3:.2f} us/step, kept {int(plan.count.sum())} cand {int(plan.cand.sum())}\n")
    value = world * B * K / (total_ms * 1e-3)
    kept = plan.count.sum().item()
    cand = plan.cand.sum().item()
    counts_on_host = int(lanes[0]["host_counts"].sum())

    # ---- multi-GPU: check the one-sided gather against an NCCL all_gather of the same result buffers ------------------
    gather_verified = None
    if gather_mode == "peer":
        ln0 = lanes[0]
        with torch.cuda.stream(ln0["stream"]):
            ln0["graphs"][0].replay() if use_graph else (ln0["post"].enqueue(*ln0["sets"][0]), ln0["tail"]())
            ln0["post"].wait_gather(0)
        torch.cuda.synchronize(dev)
        dist.barrier()
        pl0 = ln0["plan"]
        ref = torch.empty((world, pl0.packed.numel()), dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(ref, pl0.packed.clone())
        ref_rows, ref_cnt = ypb_dist.split_packed(ref, B, pl0.rows.shape[1], pl0.rows.shape[2])
        got_rows, got_cnt = ln0["post"].gathered()
        valid = (torch.arange(pl0.rows.shape[1], device=dev)[None, :] < ref_cnt[:, None]).unsqueeze(-1)
        ok = bool(torch.equal(ref_cnt, got_cnt)) and bool(torch.equal(ref_rows * valid, got_rows * valid)) and int(ref_cnt.sum()) > 0
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_verified = bool(flag.item())

    if os.environ.get("YPB_BENCH_QUICK"):  # diagnostic: only the headline timing
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": total_ms / K, "gather_verified_against_nccl": gather_verified,
                              "gather_consumer": consumer_pick, "quick": True}))
        sys.stdout.flush()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)
        os._exit(0)
    # ---- strong scaling (SURVEY 8e / north_star): ONE 64-image batch partitioned over the ranks (data/build.py:171-188) --
    strong = None
    if world > 1:
        lo, hi = ypb_dist.shard_range(B, rank, world)
        if (hi - lo) * world == B:  # equal shards: the peer ring needs the same packed size on every rank
            ssets = [([lv[lo:hi] for lv in s[0]], None) for s in sets]
            slanes = _build_lanes(dev, cfg, ssets, LANES, world, gather_mode, args.gather_lag, use_graph, lane_groups)
            sms = _time_lanes(dev, slanes, K, W, world, gather_mode, args.gather_lag, None, dev_barrier)
            strong = {"global_batch": B, "batch_per_gpu": hi - lo, "value": B * K / (sms * 1e-3), "unit": UNIT, "ms_per_step": sms / K,
                      "scaling": "strong", "note": "one 64-image batch per step split contiguously over the ranks; every rank ends "
                                                   "each step holding all 64 images' detections"}
            del slanes, ssets

    # ---- per-kernel timing with CUDA events on the launching stream (roofline of the dominant kernel) ---------------
    # (a post-processor of its own, WITHOUT the peer gather: nobody consumes - and acknowledges - ring entries here)
    post_local = _post_for(cfg, "none")
    prefix_ms = _stage_times(dev, post_local, sets, K)
    t_scan = prefix_ms[0]                  # ms: counter memset + class-scan/filter kernel
    t_decode = prefix_ms[1] - prefix_ms[0]  # ms: survivor tile box-decode kernel
    t_suppr = prefix_ms[2] - prefix_ms[1]   # ms: sort + suppress + gather kernel
    peak, peak_src = _peaks()
    scan_bytes = B * cfg.nc * cfg.anchors * esize  # the class rows: what this kernel must read
    t_scan_kernel = prefix_ms[3] if len(prefix_ms) > 3 else t_scan  # the kernel alone, no counter memset in front
    achieved = scan_bytes / (t_scan_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "scan_classes_kernel (class scan + sigmoid/confidence filter + compaction; one-wave grid, software-pipelined 128-bit loads)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": _traffic(args.dtype),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": scan_bytes,
                "algorithmic_bytes_full_head": B * in_bytes_img,
                "launch_ms": t_scan_kernel, "launch_ms_with_counter_memset": t_scan,
                "timing": "CUDA-graph replay of R back-to-back launches on rotating inputs, CUDA events on that stream; launch_ms = the "
                          "kernel alone (counters cleared once per R launches), as it runs in the real step, whose plan-owned "
                          "clean-on-exit counters need no memset node",
                "single_stream_step_ms": prefix_ms[2],
                "other_kernels_ms": {"decode_tiles_kernel": t_decode,
                                     "sort_suppress_kernel": t_suppr}}
    step_traffic = _traffic(args.dtype + "_step")
    if isinstance(step_traffic, dict):
        # the whole pipelined step against the same HBM peak: DRAM bytes of its three kernels (ncu) / the timed ms_per_step
        tot = sum(v for v in step_traffic.values() if isinstance(v, int))
        per_gpu_ms = total_ms / K
        roofline["pipelined_step"] = {"dram_bytes_per_step": tot, "achieved": tot / (per_gpu_ms * 1e-3) / 1e9,
                                      "frac": tot / (per_gpu_ms * 1e-3) / 1e9 / peak, "unit": "GB/s",
                                      "note": "all lanes overlapped: the timed region itself, per GPU"}

    # ---- dense decode kernel alone (the Detect._inference drop-in), same inputs --------------------------------------
    dense = _dense_decode_time(dev, cfg, sets, esize, peak, min(K, 50))

    # ---- the drop-in boundary: the reference's OWN call sequence (Detect._inference -> non_max_suppression,
    #      predictor.py:335-336 + detect/predict.py:54) through the wrappers patch.install() binds, on a stub module carrying
    #      Detect's attributes (head.py:70-93).  two_call = dense decode kernel + NMS-from-dense kernels; fused = the lazily
    #      decoded tensor hands the level tensors to the fused head->NMS kernels.  Wall clock per call INCLUDING the count
    #      read-back the list-of-tensors return type forces (one stream synchronisation per call, like the reference).
    dropin = _bench_dropin(dev, cfg, [s[0] for s in sets], NSETS, min(K, 200))

    # ---- end to end through the public API with HOST buffers ---------------------------------------------------------------
    e2e = _e2e(dev, cfg, sets, world, min(K, 30))

    # ---- p50 latency at B=1 through the public API (device-resident input, result counts on host) ---------------------
    one = [lv[:1].contiguous() for lv in sets[0][0]]

    def p50(fn, warm=20, reps=200):
        for _ in range(warm):
            fn()
        lat = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            lat.append((time.perf_counter() - t0) * 1e3)
        lat.sort()
        return lat[len(lat) // 2], lat[int(len(lat) * 0.9)]

    lat50, lat90 = p50(lambda: postprocess_from_head(one, cfg.strides, cfg.nc, cfg.conf, cfg.iou, max_det=cfg.max_det))
    # the same through the cached-plan serving object: static input buffers, ONE CUDA-graph launch (kernels + result
    # packing + count copy to pinned memory) and one stream synchronisation per call
    pp1 = HeadPostProcessor(cfg.nc, cfg.strides, cfg.conf, cfg.iou, max_det=cfg.max_det, max_nms=cfg.max_nms, use_graph=True)
    lat_pp50, lat_pp90 = p50(lambda: pp1(one))
    pp1e = HeadPostProcessor(cfg.nc, cfg.strides, cfg.conf, cfg.iou, max_det=cfg.max_det, max_nms=cfg.max_nms)
    lat_e50, _ = p50(lambda: pp1e(one))

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC, "batch_per_gpu": B, "global_batch": B * world,
                       "l2": f"inputs {B * in_bytes_img / 1e6:.0f} MB per step > 126 MB L2; {NSETS} rotating input sets",
                       "cuda_graph": use_graph, "lanes": LANES,
                       "timed_region": "per step: class scan, survivor decode, sort+suppress+gather (3 kernels; the plan-owned clean-on-exit counters need no memset), "
                                       "the per-image counts written to pinned HOST memory by the suppression kernel" + (", the gathered results of all ranks copied out of the peer ring" if gather_mode == "peer" else ""),
                       "parallelism": ("images sharded across ranks, no data-path collective; results gathered on every rank each step by "
                                       + {"peer": f"one-sided NVLink peer-memory stores issued by the suppression kernel into a 3-entry ring with consumer acknowledgements + an arrival-flag wait (lag {args.gather_lag} batch per lane; every gathered batch is CONSUMED - copied out - inside the step; drained before the clock stops), all inside the lane's CUDA graph",
                                          "nccl": "one packed NCCL all_gather (one communicator per lane) captured in the lane's CUDA graph",
                                          "nccl-eager": "one packed NCCL all_gather (one communicator per lane) issued from the host",
                                          "none": "NOTHING (diagnostic: independent replicas)"}[gather_mode]
                                       + ", overlapped with the other lanes") if world > 1 else "single GPU"},
            "clocks": sampler.summary(),
            "gpu_launches": (3 + (1 if gather_mode == "peer" else 0)) * K,
            "gather_verified_against_nccl": gather_verified,
            "gather_consumer": consumer_pick,
            "strong_scaling": strong,
            "e2e": e2e,
            "roofline": roofline,
            "decode_dense": dense,
            "dropin_two_call": dropin["two_call"],
            "dropin_fused": dropin["fused"],
            "latency_b1_ms_p50": lat50,
            "latency_b1_ms_p90": lat90,
            "latency_b1_ms_p50_cached_plan": lat_pp50,
            "latency_b1_ms_p90_cached_plan": lat_pp90,
            "latency_b1_ms_p50_cached_plan_no_graph": lat_e50,
            "detections_last_step": {"kept": int(kept), "candidates": int(cand), "counts_on_host": counts_on_host},
        }
    if world == 1 and not args.no_extra:
        # ---- the other BASELINE.json configs and the 16-bit head, measured by the same machinery (driver-run numbers) ------
        del lanes
        extra = {}
        for name in ("c3_val_stress_b32", "c4_p6_1280_b16", "c5_obb_1024_b16"):
            extra[name] = _measure_config(dev, name, dtype, min(K, 200))
        line["configs"] = extra
        # ---- the persistent TMA-fed form of the class scan (scan_kernel="tma"), same workload: faster as a single kernel,
        #      slower in the multi-lane pipeline (its ring fills the SM's shared memory, so other lanes' CTAs cannot co-reside)
        def tma_record(tsets, esz):
            tl = _build_lanes(dev, cfg, tsets, LANES, 1, "none", 0, True, None, "tma")
            tms = _time_lanes(dev, tl, K, W, 1, "none", 0)
            tpre = _stage_times(dev, tl[0]["post"], tsets, K)
            tk = tpre[3] if len(tpre) > 3 else tpre[0]
            sb = B * cfg.nc * cfg.anchors * esz
            pp_t = HeadPostProcessor(cfg.nc, cfg.strides, cfg.conf, cfg.iou, max_det=cfg.max_det, max_nms=cfg.max_nms, use_graph=True,
                                     scan_kernel="tma")
            one_t = [lv[:1].contiguous() for lv in tsets[0][0]]
            l50, _ = p50(lambda: pp_t(one_t))
            return {"kernel": "scan_classes_tma_kernel (cp.async.bulk.tensor.3d ring, 1 CTA per SM)", "launch_ms": tk,
                    "achieved": sb / (tk * 1e-3) / 1e9, "frac": sb / (tk * 1e-3) / 1e9 / peak, "unit": "GB/s", "peak": peak,
                    "single_stream_step_ms": tpre[2], "pipelined_value": B * K / (tms * 1e-3), "pipelined_ms_per_step": tms / K,
                    "latency_b1_ms_p50_cached_plan": l50}

        line["scan_tma"] = {"f32" if args.dtype == "f32" else args.dtype: tma_record(sets, esize)}
        if args.dtype == "f32":
            sets16 = [([lv.to(torch.bfloat16) for lv in s[0]], None) for s in sets]
            del sets
            torch.cuda.empty_cache()
            l16 = _build_lanes(dev, cfg, sets16, LANES, 1, "none", 0, True)
            ms16 = _time_lanes(dev, l16, K, W, 1, "none", 0)
            pre16 = _stage_times(dev, l16[0]["post"], sets16, K)
            sb16 = B * cfg.nc * cfg.anchors * 2
            t16 = pre16[3] if len(pre16) > 3 else pre16[0]
            line["bf16"] = {"value": B * K / (ms16 * 1e-3), "unit": UNIT, "ms_per_step": ms16 / K,
                            "roofline": {"kernel": "scan_classes_kernel", "achieved": sb16 / (t16 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                         "frac": sb16 / (t16 * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": sb16,
                                         "launch_ms": t16, "launch_ms_with_counter_memset": pre16[0], "traffic": _traffic("bf16"), "single_stream_step_ms": pre16[2],
                                         "other_kernels_ms": {"decode_tiles_kernel": pre16[1] - pre16[0], "sort_suppress_kernel": pre16[2] - pre16[1]}},
                            "decode_dense": _dense_decode_time(dev, cfg, sets16, 2, peak, min(K, 50)),
                            "e2e": _e2e(dev, cfg, sets16, 1, min(K, 30))}
            del l16
            line["scan_tma"]["bf16"] = tma_record(sets16, 2)
            sets = None
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(cfg)
        print(json.dumps(line))
    if world > 1:
        # leave without tearing the communicators down: destroy_process_group() was seen to hang with CUDA graphs that
        # hold captured collectives; every rank has finished its device work here
        sys.stdout.flush()
        dist.barrier()
        torch.cuda.synchronize(dev)
        os._exit(0)
    return 0


# ----------------------------------------------------------------------------------------------------------------
# the drop-in boundary
# ----------------------------------------------------------------------------------------------------------------
class _StubDetect:
    """Attribute surface of the reference's ``Detect`` (head.py:70-93) without the convolutions; the methods bound below
    are the wrappers ``patch.install()`` puts on the reference's classes."""

    dynamic = export = end2end = xyxy = training = False
    format = None
    shape = None
    max_det = 300

    def __init__(self, cfg, dev):
        self.nc, self.nl, self.reg_max, self.no = cfg.nc, len(cfg.strides), cfg.reg_max, cfg.no
        self.stride = torch.tensor([float(s) for s in cfg.strides], device=dev)
        self.anchors = self.strides = torch.empty(0)


def _bench_dropin(dev, cfg, sets, nsets, reps):
    import types

    from ultralytics_pro_b200 import head, lazy, nms, patch

    def boom(*a, **k):
        raise RuntimeError("reference function called on the GPU path")

    mod = _StubDetect(cfg, dev)
    mod._inference = types.MethodType(patch._wrap_inference(boom, head.detect_inference), mod)
    nms_fn = patch._wrap_nms(boom, nms.non_max_suppression)
    out = {}
    B = sets[0][0].shape[0]
    for name, lazy_on in (("two_call", False), ("fused", True)):
        lazy.ENABLED = lazy_on

        def call(i):
            with torch.inference_mode():
                y = mod._inference(sets[i % nsets])
                return nms_fn((y, sets[i % nsets]), cfg.conf, cfg.iou, max_det=cfg.max_det)

        for i in range(5):
            res = call(i)
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for i in range(reps):
            res = call(i)
        b.record()
        torch.cuda.synchronize(dev)
        wall = (time.perf_counter() - t0) / reps
        out[name] = {"imgs_per_s": B / wall, "ms_per_call": wall * 1e3, "ms_per_call_cuda_events": a.elapsed_time(b) / reps,
                     "calls": reps, "kept_last_call": int(sum(r.shape[0] for r in res)),
                     "path": ("LazyDecoded -> ypb_nms_from_head (scan_classes + decode_tiles + sort_suppress)" if lazy_on
                              else "ypb_decode_dense + ypb_nms_from_dense (filter_from_dense + sort_suppress)"),
                     "includes": "python + ctypes + kernels + the count D2H and stream sync of every call"}
    lazy.ENABLED = False
    return out


# ----------------------------------------------------------------------------------------------------------------
# --post-rows: measurement of the SURVEY.md 8f rows
# ----------------------------------------------------------------------------------------------------------------
def _gpu_ms(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def _cpu_ms(fn, budget=3.0):
    fn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget:
        fn()
        n += 1
    return (time.perf_counter() - t0) / n * 1e3


def run_post_rows(args):
    """--post-rows: the SURVEY.md 8f rows (the steps after NMS), one JSON line per kernel: CUDA-event time, the roofline it is
    bound by, and - as that row's cpu_baseline - the oracle timed on the host CPU on a bounded sample."""
    import numpy as np  # noqa: F401

    from oracle import result_ops_oracle as ro  # cpu_baseline leg of each row
    from ultralytics_pro_b200 import ops, val
    from ultralytics_pro_b200.head import decode_head, decode_keypoints, postprocess_from_head
    from ultralytics_pro_b200.nms import non_max_suppression
    from ultralytics_pro_b200.synth import CONFIGS, make_head_batch

    dev = torch.device("cuda:0")
    peak, src = _peaks()
    cfg = CONFIGS["c2_v8x_640_b64"]
    B, A = 64, cfg.anchors
    g = torch.Generator().manual_seed(5)
    out = []

    # ---- 1. construct_result rescale of a batch's kept rows ----------------------------------------------------------
    rows = (torch.rand(B, 300, 6, generator=g) * 640).to(dev)
    cnt = torch.full((B,), 100, dtype=torch.int32, device=dev)
    shapes = [(480 + 8 * i, 640 + 4 * i, 3) for i in range(B)]
    ops.scale_results(rows, cnt, (640, 640), shapes)  # builds the transform array
    ms = _gpu_ms(lambda: ops.scale_results(rows, cnt, (640, 640), shapes))
    r_cpu = rows[0, :100, :4].cpu().numpy()
    c = _cpu_ms(lambda: ro.scale_boxes_oracle((640, 640), r_cpu, shapes[0]), 1.0) * B
    out.append({"row": "8f-1 scale_boxes of a batch's kept rows (B=64 x 100 rows)", "kernel": "scale_rows_kernel", "ms": ms,
                "bound": "launch latency (one launch; 150 KB touched)", "cpu_oracle_ms": c, "cpu_sample": "numpy oracle, 1 image x 64"})

    # ---- 2. Pose.kpts_decode, dense ------------------------------------------------------------------------------------
    kp = torch.randn(B, 51, A, generator=g).to(dev)
    ms = _gpu_ms(lambda: decode_keypoints(kp, cfg.level_hw, cfg.strides, (17, 3)), reps=30)
    nbytes = 2 * kp.numel() * 4
    kp_cpu = kp[:4].cpu()
    c = _cpu_ms(lambda: ro.kpts_decode_oracle(kp_cpu, cfg.level_hw, cfg.strides, (17, 3)), 2.0) * (B / 4)
    out.append({"row": "8f-2 Pose.kpts_decode dense (B=64, 17x3 keypoints, 8400 anchors, fp32)", "kernel": "kpts_decode_kernel", "ms": ms,
                "bound": "hbm", "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / ms / 1e6, "peak_gbs": peak, "peak_source": src,
                "frac": nbytes / ms / 1e6 / peak, "cpu_oracle_ms": c, "cpu_sample": "torch CPU oracle on 4 images x 16"})

    # ---- 3. pose post-processing: dense decode + cat + NMS vs fused riders -----------------------------------------------
    pcfg = cfg.__class__("pose", 640, (8, 16, 32), 1, B)
    levels = [lv.to(dev) for lv in make_head_batch(pcfg, batch=B, seed=3)[0]]

    def two_call():
        dense = torch.cat([decode_head(levels, pcfg.strides, 1), decode_keypoints(kp, pcfg.level_hw, pcfg.strides, (17, 3))], 1)
        return non_max_suppression(dense, 0.25, 0.7, nc=1)

    def fused():
        return postprocess_from_head(levels, pcfg.strides, 1, 0.25, 0.7, kpt_logits=kp, kpt_shape=(17, 3))

    out.append({"row": "8f-2 Pose post-process B=64 (boxes + 17x3 keypoints -> kept rows)", "two_call_ms": _gpu_ms(two_call, 20),
                "fused_riders_ms": _gpu_ms(fused, 20), "kept_rows": int(sum(t.shape[0] for t in fused())),
                "note": "fused = postprocess_from_head(kpt_logits=...): keypoints decoded for kept anchors only; both include the count D2H sync"})

    # ---- 4. process_mask, batched -----------------------------------------------------------------------------------------
    Bm, n_img = 16, 100
    protos = torch.randn(Bm, 32, 160, 160, generator=g).to(dev)
    mrows = torch.zeros(Bm, 300, 38)
    xy = torch.rand(Bm, 300, 2, generator=g) * 500
    mrows[..., :2], mrows[..., 2:4] = xy, xy + torch.rand(Bm, 300, 2, generator=g) * 200 + 10
    mrows[..., 6:] = torch.randn(Bm, 300, 32, generator=g)
    mrows = mrows.to(dev)
    counts = [n_img] * Bm
    ms = _gpu_ms(lambda: ops.process_masks_batched(protos, mrows, counts, (640, 640), True), reps=10, warm=2)
    nbytes = Bm * n_img * 640 * 640
    p_cpu, r_cpu2 = protos[0].cpu(), mrows[0, :n_img].cpu()
    c = _cpu_ms(lambda: ro.process_mask_oracle(p_cpu, r_cpu2[:, 6:], r_cpu2[:, :4], (640, 640), True), 4.0) * Bm
    out.append({"row": "8f-2 process_mask(upsample) B=16 x 100 detections -> (1600, 640, 640) uint8", "kernel": "process_mask_kernel", "ms": ms,
                "bound": "hbm (write of the uint8 masks)", "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / ms / 1e6, "peak_gbs": peak,
                "peak_source": src, "frac": nbytes / ms / 1e6 / peak, "cpu_oracle_ms": c, "cpu_sample": "torch CPU oracle, 1 image x 16"})

    # ---- 5. validator matching -----------------------------------------------------------------------------------------------
    M = 40
    gxy = torch.rand(B, M, 2, generator=g) * 500
    gt = torch.cat([gxy, gxy + torch.rand(B, M, 2, generator=g) * 100 + 8], 2)
    gcls = torch.randint(0, 80, (B, M), generator=g).float()
    src_i = torch.randint(0, M, (B, 300), generator=g)
    pr = torch.gather(gt, 1, src_i[..., None].expand(-1, -1, 4)) + torch.randn(B, 300, 4, generator=g) * 4
    vrows = torch.cat([pr, torch.rand(B, 300, 1, generator=g), torch.gather(gcls, 1, src_i)[..., None]], 2).to(dev)
    labels = torch.cat([gcls[..., None], gt], 2).reshape(-1, 5).to(dev)
    vc = torch.full((B,), 300, dtype=torch.int32, device=dev)
    iouv = torch.linspace(0.5, 0.95, 10).tolist()
    ms = _gpu_ms(lambda: val.match_batch(iouv, vrows, vc, labels, [M] * B))
    pb, pc_, gb_, gc_ = pr[0].numpy(), vrows[0, :, 5].cpu().numpy(), gt[0].numpy(), gcls[0].numpy()
    c = _cpu_ms(lambda: ro.match_predictions_oracle(pc_, gc_, ro.box_iou_oracle(gb_, pb), iouv), 2.0) * B
    out.append({"row": "8f-4 box_iou + match_predictions, B=64 x 300 detections x 40 labels x 10 IoU levels", "kernel": "match_predictions_kernel",
                "ms": ms, "bound": "latency (one CTA per image)", "pairs": B * 300 * M, "cpu_oracle_ms": c, "cpu_sample": "numpy oracle, 1 image x 64"})

    # ---- 6. exporter NMSModel flavour and the end2end top-k, on the decoded C2 batch ---------------------------------------
    from ultralytics_pro_b200.export_nms import nms_model_postprocess
    from ultralytics_pro_b200.head import detect_postprocess

    dl = [lv.to(dev) for lv in make_head_batch(cfg, batch=B, seed=9)[0]]
    y_xyxy = decode_head(dl, cfg.strides, cfg.nc, xyxy=True)
    ms = _gpu_ms(lambda: nms_model_postprocess(y_xyxy, (640, 640), cfg.nc, 0.25, 0.45, 300), reps=30)
    y_cpu = y_xyxy[:4].cpu()
    c = _cpu_ms(lambda: ro.nms_model_oracle(y_cpu, (640, 640), cfg.nc, 0.25, 0.45, 300), 3.0) * (B / 4)
    out.append({"row": "8f-3 exporter NMSModel post-processing, B=64 x 8400 anchors x 80 classes -> (64, 300, 6) padded, no host sync",
                "kernels": "filter_from_dense_kernel + sort_suppress_kernel (normalised-offset mode)", "ms": ms,
                "bound": "hbm (one pass over the 172 MB of scores) + latency", "achieved_gbs": B * 80 * A * 4 / ms / 1e6,
                "cpu_oracle_ms": c, "cpu_sample": "torch CPU oracle (torchvision nms), 4 images x 16"})
    preds = y_xyxy.permute(0, 2, 1)
    ms = _gpu_ms(lambda: detect_postprocess(preds, 300, cfg.nc), reps=20)
    p_cpu = preds[:4].cpu()
    c = _cpu_ms(lambda: ro.detect_postprocess_oracle(p_cpu, 300, cfg.nc), 3.0) * (B / 4)
    out.append({"row": "8f-4 Detect.postprocess end2end top-k, B=64 x 8400 anchors x 80 classes -> (64, 300, 6)",
                "kernels": "filter_from_dense_kernel + sort_suppress_kernel over all anchors, then the same two on the K kept anchors",
                "bound": "host (two plans + ~10 small torch ops per call) above one hbm pass over the 172 MB of scores", "ms": ms,
                "achieved_gbs": B * 80 * A * 4 / ms / 1e6, "cpu_oracle_ms": c, "cpu_sample": "torch CPU oracle (one stable sort), 4 images x 16"})

    for o in out:
        print(json.dumps(o))
    return 0



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16", "f16"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl", "nccl-eager", "none"],
                    help="multi-GPU result gather: peer = one-sided stores over NVLink peer memory from the suppression kernel "
                         "(default); nccl = one all_gather per step captured in each lane's CUDA graph (one communicator per "
                         "lane); nccl-eager = the same issued from the host")
    ap.add_argument("--gather-lag", type=int, default=1,
                    help="peer gather: each step waits for the arrival of the results of the lane's batch this many batches back "
                         "(0 = the batch just processed); the lanes are drained with lag 0 before the timed region ends")
    ap.add_argument("--lanes", type=int, default=5, help="independent pipelines (streams) the steps are spread over")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the C3/C4/C5 and bf16 sub-records (N=1 only)")
    ap.add_argument("--post-rows", action="store_true",
                    help="instead of the headline line, print one JSON line per SURVEY 8f row (rescale, keypoints, masks, matching, "
                         "NMSModel, top-k) with its CUDA-event time, roofline and CPU-oracle baseline")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.post_rows:
        return run_post_rows(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
