"""Diagnostic: device-resident timing of the fused path on every BASELINE.json config (not the bench line)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultralytics_pro_b200.pipeline import HeadPostProcessor
from ultralytics_pro_b200.synth import CONFIGS, make_head_batch
from ultralytics_pro_b200.head import decode_head

dev = torch.device("cuda:0")
out = {}
for name, cfg in CONFIGS.items():
    for dtype in (torch.float32, torch.bfloat16):
        lv, ang = make_head_batch(cfg, seed=1000, device=dev, dtype=dtype)
        post = HeadPostProcessor(cfg.nc, cfg.strides, cfg.conf, cfg.iou, multi_label=cfg.multi_label,
                                 agnostic=cfg.agnostic, rotated=cfg.rotated, max_det=cfg.max_det, max_nms=cfg.max_nms)
        for _ in range(3):
            plan = post.enqueue(lv, ang)
        torch.cuda.synchronize()
        n = 20
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n)]
        for i in range(n):
            for j, st in enumerate((1, 2, 4)):
                ev[i][j].record(); post.enqueue(lv, ang, stage=st)
            ev[i][3].record()
        torch.cuda.synchronize()
        t = [sum(e[j].elapsed_time(e[j + 1]) for e in ev) / n for j in range(3)]
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            post.enqueue(lv, ang)
        b.record(); torch.cuda.synchronize()
        total = a.elapsed_time(b) / n
        if cfg.rotated:
            f = lambda: decode_head(lv, cfg.strides, cfg.nc, angle=ang, angle_is_logit=True, append_angle=True)
        else:
            f = lambda: decode_head(lv, cfg.strides, cfg.nc)
        for _ in range(3): f()
        a.record()
        for i in range(n): f()
        b.record(); torch.cuda.synchronize()
        dense = a.elapsed_time(b) / n
        es = 4 if dtype == torch.float32 else 2
        dense_bytes = cfg.batch * cfg.anchors * ((cfg.no + (1 if cfg.rotated else 0)) + (4 + cfg.nc + (1 if cfg.rotated else 0))) * es
        rec = {"batch": cfg.batch, "dtype": str(dtype).split(".")[-1], "scan_ms": round(t[0], 4), "decode_tiles_ms": round(t[1], 4),
               "sort_suppress_ms": round(t[2], 4), "fused_total_ms": round(total, 4), "imgs_per_s": round(cfg.batch / total * 1e3),
               "dense_decode_ms": round(dense, 4), "dense_GBps": round(dense_bytes / dense / 1e6),
               "cand_per_img": int(plan.cand.float().mean().item()), "kept_per_img": int(plan.count.float().mean().item())}
        out[f"{name}/{rec['dtype']}"] = rec
        print(name, json.dumps(rec), flush=True)
json.dump(out, open("gpurun_out/config_sweep.json", "w"), indent=1)
