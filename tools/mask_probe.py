"""Diagnostic: where process_mask's time goes - plain memset vs the kernel with boxes nobody can see vs the real workload."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultralytics_pro_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
Bm, n = 16, 100
protos = torch.randn(Bm, 32, 160, 160, generator=g).to(dev)
rows = torch.zeros(Bm, 300, 38)
xy = torch.rand(Bm, 300, 2, generator=g) * 500
rows[..., :2], rows[..., 2:4] = xy, xy + torch.rand(Bm, 300, 2, generator=g) * 200 + 10
rows[..., 6:] = torch.randn(Bm, 300, 32, generator=g)
rows = rows.to(dev)
none = rows.clone(); none[..., :4] = torch.tensor([-50.0, -50.0, -40.0, -40.0], device=dev)   # boxes outside the image
full = rows.clone(); full[..., :4] = torch.tensor([0.0, 0.0, 640.0, 640.0], device=dev)        # boxes = the whole image
counts = [n] * Bm
buf = torch.empty(Bm * n, 640, 640, dtype=torch.uint8, device=dev)
def t(fn, reps=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
res = {"memset_655MB_ms": t(lambda: buf.zero_()),
       "fill1_655MB_ms": t(lambda: buf.fill_(1)),
       "kernel_no_box_visible_ms": t(lambda: ops.process_masks_batched(protos, none, counts, (640, 640), True)),
       "kernel_real_boxes_ms": t(lambda: ops.process_masks_batched(protos, rows, counts, (640, 640), True)),
       "kernel_full_image_boxes_ms": t(lambda: ops.process_masks_batched(protos, full, counts, (640, 640), True))}
print(json.dumps(res))
