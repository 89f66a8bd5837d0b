"""Diagnostic: B=1 (or argv[1]) p50 latency of the fused path through the public entry points, C2 shapes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultralytics_pro_b200.head import postprocess_from_head
from ultralytics_pro_b200.pipeline import HeadPostProcessor
from ultralytics_pro_b200.synth import CONFIGS, make_head_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = CONFIGS["c2_v8x_640_b64"]
dev = torch.device("cuda:0")
levels = [lv[:B].contiguous() for lv in make_head_batch(cfg, batch=max(B, 1), seed=1000, device=dev)[0]]


REPS = int(os.environ.get('YPB_LAT_REPS', '300'))


def p50(fn, warm=30, reps=None):
    reps = reps or REPS
    warm = min(warm, reps)
    for _ in range(warm):
        fn()
    lat = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        lat.append((time.perf_counter() - t0) * 1e6)
    lat.sort()
    return round(lat[len(lat) // 2], 1), round(lat[int(len(lat) * 0.9)], 1)


out = {"batch": B, "env": {k: v for k, v in os.environ.items() if k.startswith("YPB_")}}
out["postprocess_from_head_us"] = p50(lambda: postprocess_from_head(levels, cfg.strides, cfg.nc, cfg.conf, cfg.iou, max_det=cfg.max_det))
pp = HeadPostProcessor(cfg.nc, cfg.strides, cfg.conf, cfg.iou, max_det=cfg.max_det, max_nms=cfg.max_nms, use_graph=True)
out["cached_plan_graph_us"] = p50(lambda: pp(levels))
ppe = HeadPostProcessor(cfg.nc, cfg.strides, cfg.conf, cfg.iou, max_det=cfg.max_det, max_nms=cfg.max_nms)
out["cached_plan_no_graph_us"] = p50(lambda: ppe(levels))
# GPU-side time of the graph alone (CUDA events around back-to-back replays)
graph = next(iter(pp._graphs.values()))[0]
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
a.record()
for _ in range(50):
    graph.replay()
b.record()
torch.cuda.synchronize()
out["graph_replay_gpu_us"] = round(a.elapsed_time(b) * 1000 / 50, 1)
# isolated latency of each prefix of the step (idle GPU, one graph replay between two events, synchronised every time)
st = torch.cuda.Stream(dev)
pref = {}
with torch.cuda.stream(st):
    for mask in (1, 3, 7):
        ppe.enqueue(levels, stage=mask)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            ppe.enqueue(levels, stage=mask)
        ts = []
        for _ in range(60):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); g.replay(); e1.record(st); st.synchronize()
            ts.append(e0.elapsed_time(e1) * 1000)
        ts.sort()
        pref[mask] = round(ts[len(ts) // 2], 1)
    ts = []
    for _ in range(60):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); graph.replay(); e1.record(st); st.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000)
    ts.sort()
    pref["full_graph"] = round(ts[len(ts) // 2], 1)
out["isolated_prefix_us"] = pref
print(json.dumps(out))
