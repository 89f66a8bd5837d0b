#!/usr/bin/env python
"""Measurement of the SURVEY.md 8f rows (the steps after NMS): one JSON line per kernel with its CUDA-event time, the
roofline it is bound by, and the oracle timed on the host CPU on a bounded sample.  Not the headline bench (bench.py).

    python tools/bench_post.py > gpurun_out/post_rows.jsonl
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import result_ops_oracle as ro  # noqa: E402  (cpu_baseline leg only)
from ultralytics_pro_b200 import ops, val  # noqa: E402
from ultralytics_pro_b200.head import decode_head, decode_keypoints, postprocess_from_head  # noqa: E402
from ultralytics_pro_b200.nms import non_max_suppression  # noqa: E402
from ultralytics_pro_b200.synth import CONFIGS, make_head_batch  # noqa: E402


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def gpu_ms(fn, reps=50, warm=5):
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


def cpu_ms(fn, budget=3.0):
    fn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget:
        fn()
        n += 1
    return (time.perf_counter() - t0) / n * 1e3


def main():
    dev = torch.device("cuda:0")
    peak, src = peak_gbs()
    cfg = CONFIGS["c2_v8x_640_b64"]
    B, A = 64, cfg.anchors
    g = torch.Generator().manual_seed(5)
    out = []

    # ---- 1. construct_result rescale of a batch's kept rows ----------------------------------------------------------
    rows = (torch.rand(B, 300, 6, generator=g) * 640).to(dev)
    cnt = torch.full((B,), 100, dtype=torch.int32, device=dev)
    shapes = [(480 + 8 * i, 640 + 4 * i, 3) for i in range(B)]
    ops.scale_results(rows, cnt, (640, 640), shapes)  # builds the transform array
    ms = gpu_ms(lambda: ops.scale_results(rows, cnt, (640, 640), shapes))
    r_cpu = rows[0, :100, :4].cpu().numpy()
    c = cpu_ms(lambda: ro.scale_boxes_oracle((640, 640), r_cpu, shapes[0]), 1.0) * B
    out.append({"row": "8f-1 scale_boxes of a batch's kept rows (B=64 x 100 rows)", "kernel": "scale_rows_kernel", "ms": ms,
                "bound": "launch latency (one launch; 150 KB touched)", "cpu_oracle_ms": c, "cpu_sample": "numpy oracle, 1 image x 64"})

    # ---- 2. Pose.kpts_decode, dense ------------------------------------------------------------------------------------
    kp = torch.randn(B, 51, A, generator=g).to(dev)
    ms = gpu_ms(lambda: decode_keypoints(kp, cfg.level_hw, cfg.strides, (17, 3)), reps=30)
    nbytes = 2 * kp.numel() * 4
    kp_cpu = kp[:4].cpu()
    c = cpu_ms(lambda: ro.kpts_decode_oracle(kp_cpu, cfg.level_hw, cfg.strides, (17, 3)), 2.0) * (B / 4)
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

    out.append({"row": "8f-2 Pose post-process B=64 (boxes + 17x3 keypoints -> kept rows)", "two_call_ms": gpu_ms(two_call, 20),
                "fused_riders_ms": gpu_ms(fused, 20), "kept_rows": int(sum(t.shape[0] for t in fused())),
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
    ms = gpu_ms(lambda: ops.process_masks_batched(protos, mrows, counts, (640, 640), True), reps=10, warm=2)
    nbytes = Bm * n_img * 640 * 640
    p_cpu, r_cpu2 = protos[0].cpu(), mrows[0, :n_img].cpu()
    c = cpu_ms(lambda: ro.process_mask_oracle(p_cpu, r_cpu2[:, 6:], r_cpu2[:, :4], (640, 640), True), 4.0) * Bm
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
    ms = gpu_ms(lambda: val.match_batch(iouv, vrows, vc, labels, [M] * B))
    pb, pc_, gb_, gc_ = pr[0].numpy(), vrows[0, :, 5].cpu().numpy(), gt[0].numpy(), gcls[0].numpy()
    c = cpu_ms(lambda: ro.match_predictions_oracle(pc_, gc_, ro.box_iou_oracle(gb_, pb), iouv), 2.0) * B
    out.append({"row": "8f-4 box_iou + match_predictions, B=64 x 300 detections x 40 labels x 10 IoU levels", "kernel": "match_predictions_kernel",
                "ms": ms, "bound": "latency (one CTA per image)", "pairs": B * 300 * M, "cpu_oracle_ms": c, "cpu_sample": "numpy oracle, 1 image x 64"})

    # ---- 6. exporter NMSModel flavour and the end2end top-k, on the decoded C2 batch ---------------------------------------
    from ultralytics_pro_b200.export_nms import nms_model_postprocess
    from ultralytics_pro_b200.head import detect_postprocess

    dl = [lv.to(dev) for lv in make_head_batch(cfg, batch=B, seed=9)[0]]
    y_xyxy = decode_head(dl, cfg.strides, cfg.nc, xyxy=True)
    ms = gpu_ms(lambda: nms_model_postprocess(y_xyxy, (640, 640), cfg.nc, 0.25, 0.45, 300), reps=30)
    y_cpu = y_xyxy[:4].cpu()
    c = cpu_ms(lambda: ro.nms_model_oracle(y_cpu, (640, 640), cfg.nc, 0.25, 0.45, 300), 3.0) * (B / 4)
    out.append({"row": "8f-3 exporter NMSModel post-processing, B=64 x 8400 anchors x 80 classes -> (64, 300, 6) padded, no host sync",
                "kernels": "filter_from_dense_kernel + sort_suppress_kernel (normalised-offset mode)", "ms": ms,
                "bound": "hbm (one pass over the 172 MB of scores) + latency", "achieved_gbs": B * 80 * A * 4 / ms / 1e6,
                "cpu_oracle_ms": c, "cpu_sample": "torch CPU oracle (torchvision nms), 4 images x 16"})
    preds = y_xyxy.permute(0, 2, 1)
    ms = gpu_ms(lambda: detect_postprocess(preds, 300, cfg.nc), reps=20)
    p_cpu = preds[:4].cpu()
    c = cpu_ms(lambda: ro.detect_postprocess_oracle(p_cpu, 300, cfg.nc), 3.0) * (B / 4)
    out.append({"row": "8f-4 Detect.postprocess end2end top-k, B=64 x 8400 anchors x 80 classes -> (64, 300, 6)",
                "kernels": "2 x (filter_from_dense_kernel + sort_suppress_kernel)", "ms": ms,
                "bound": "hbm (two passes over the 172 MB of scores) + the per-image radix sort of 8400 keys",
                "achieved_gbs": 2 * B * 80 * A * 4 / ms / 1e6, "cpu_oracle_ms": c, "cpu_sample": "torch CPU oracle (one stable sort), 4 images x 16"})

    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
