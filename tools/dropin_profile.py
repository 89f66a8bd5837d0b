"""Diagnostic: where the host time of one drop-in call (Detect._inference -> non_max_suppression) goes."""
import cProfile, pstats, sys, time, types, io
sys.path.insert(0, ".")
import torch
from bench import _StubDetect
from ultralytics_pro_b200 import head, lazy, nms, patch
from ultralytics_pro_b200.synth import CONFIGS, make_head_batch

dev = torch.device("cuda:0")
cfg = CONFIGS["c2_v8x_640_b64"]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sets = [[lv.to(dev) for lv in make_head_batch(cfg, batch=B, seed=s)[0]] for s in range(3)]
mod = _StubDetect(cfg, dev)
mod._inference = types.MethodType(patch._wrap_inference(None, head.detect_inference), mod)
nms_fn = patch._wrap_nms(None, nms.non_max_suppression)
for lazy_on in (False, True):
    lazy.ENABLED = lazy_on
    def call(i):
        with torch.inference_mode():
            y = mod._inference(sets[i % 3])
            return nms_fn((y, sets[i % 3]), cfg.conf, cfg.iou, max_det=cfg.max_det)
    for i in range(10): call(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(300): call(i)
    torch.cuda.synchronize()
    print(f"lazy={lazy_on} B={B}: {(time.perf_counter()-t0)/300*1e6:.1f} us/call")
    pr = cProfile.Profile(); pr.enable()
    for i in range(300): call(i)
    pr.disable()
    st = io.StringIO(); pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(22); print(st.getvalue()[:5000])
