"""Diagnostic: per-phase clock64 marks of the sort+suppress kernel on the bench workload (not part of the product)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultralytics_pro_b200 import _cabi
from ultralytics_pro_b200.pipeline import HeadPostProcessor
from ultralytics_pro_b200.synth import CONFIGS, make_head_batch

cfg = CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2_v8x_640_b64"]
dev = torch.device("cuda:0")
B = cfg.batch
lv, ang = make_head_batch(cfg, batch=B, seed=1000, device=dev)
post = HeadPostProcessor(cfg.nc, cfg.strides, cfg.conf, cfg.iou, multi_label=cfg.multi_label, rotated=cfg.rotated)
for _ in range(3):
    post.enqueue(lv, ang)
torch.cuda.synchronize()
buf = torch.zeros(B * 32, dtype=torch.int64, device=dev)
lib = _cabi.load()
lib.ypb_debug_set_phase_buffer(buf.data_ptr())
plan = post.enqueue(lv, ang)
torch.cuda.synchronize()
lib.ypb_debug_set_phase_buffer(None)
m = buf.view(B, 32).cpu()
cand = plan.cand.cpu(); cnt = plan.count.cpu()
for b in list(range(min(B, 6))):
    r = m[b]
    marks = [(i, int(r[i] - r[0]) if i not in (23,24,25,26) else int(r[i])) for i in range(32) if r[i] != 0]
    print(f"img {b}: cand {int(cand[b])} kept {int(cnt[b])} cycles:", marks)
tot = (m[:, 31] - m[:, 0]).float()
print("total cycles per CTA: mean %.0f max %.0f min %.0f" % (tot.mean(), tot.max(), tot.min()))
print("rank phase mean %.0f" % (m[:, 1] - m[:, 0]).float().mean())
