"""Diagnostic: dense decode (Detect._inference drop-in) time for 16-bit heads, C2 shapes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultralytics_pro_b200.head import decode_head
from ultralytics_pro_b200.synth import CONFIGS, make_head_batch
cfg = CONFIGS["c2_v8x_640_b64"]; dev = torch.device("cuda:0")
for dt in (torch.bfloat16, torch.float16, torch.float32):
    sets = [[lv.to(dev) for lv in make_head_batch(cfg, batch=64, seed=s, dtype=dt)[0]] for s in range(3)]
    for i in range(5): decode_head(sets[i % 3], cfg.strides, cfg.nc)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(60): decode_head(sets[i % 3], cfg.strides, cfg.nc)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 60
    es = 4 if dt == torch.float32 else 2
    nb = 64 * (144 + 84) * 8400 * es
    print(json.dumps({"dtype": str(dt), "vec16": os.environ.get("YPB_DENSE16_VEC", "4"), "ms": ms, "gbs": nb / ms / 1e6, "frac": nb / ms / 1e6 / 6550.4}))
