"""Diagnostic: single-stream stage times of the fused path (graph replays, rotating inputs) + a fused == two-call check."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import _stage_times
from ultralytics_pro_b200.pipeline import HeadPostProcessor
from ultralytics_pro_b200.head import decode_head, postprocess_from_head
from ultralytics_pro_b200.nms import non_max_suppression
from ultralytics_pro_b200.synth import CONFIGS, make_head_batch

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "c2_v8x_640_b64"
dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[sys.argv[2] if len(sys.argv) > 2 else "f32"]
cfg = CONFIGS[name]
sets = [make_head_batch(cfg, seed=1000 + s, device=dev, dtype=dtype) for s in range(6)]
post = HeadPostProcessor(cfg.nc, cfg.strides, cfg.conf, cfg.iou, multi_label=cfg.multi_label, agnostic=cfg.agnostic,
                         rotated=cfg.rotated, max_det=cfg.max_det, max_nms=cfg.max_nms)
pre = _stage_times(dev, post, sets, 200)
ok = None
if not cfg.rotated:
    lv = sets[0][0]
    one = postprocess_from_head(lv, cfg.strides, cfg.nc, cfg.conf, cfg.iou, multi_label=cfg.multi_label)
    two = non_max_suppression(decode_head(lv, cfg.strides, cfg.nc), cfg.conf, cfg.iou, multi_label=cfg.multi_label)
    ok = all(torch.equal(a, b) for a, b in zip(one, two))
es = 4 if dtype == torch.float32 else 2
sb = cfg.batch * cfg.nc * cfg.anchors * es
print(json.dumps({"cfg": name, "dtype": str(dtype), "env": {k: v for k, v in os.environ.items() if k.startswith("YPB_")},
                  "scan_us": round(pre[0] * 1e3, 2), "scan_kernel_only_us": round(pre[3] * 1e3, 2) if len(pre) > 3 else None, "frac_kernel_only": round(sb / pre[3] / 1e6 / 6550.4, 3) if len(pre) > 3 else None, "decode_us": round((pre[1] - pre[0]) * 1e3, 2), "suppress_us": round((pre[2] - pre[1]) * 1e3, 2),
                  "scan_gbs": round(sb / pre[0] / 1e6), "frac": round(sb / pre[0] / 1e6 / 6550.4, 3), "fused_eq_two_call": ok}))
