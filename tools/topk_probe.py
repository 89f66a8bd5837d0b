"""Diagnostic: end2end top-k (Detect.postprocess drop-in) - wall time per call against the GPU time of the same call replayed
as a CUDA graph (no host work), B=64 x 8400 anchors x 80 classes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultralytics_pro_b200.head import detect_postprocess

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
preds = torch.cat([torch.rand(64, 8400, 4, generator=g) * 640, torch.rand(64, 8400, 80, generator=g)], 2).to(dev)
for _ in range(5):
    out = detect_postprocess(preds, 300, 80)
torch.cuda.synchronize()
REPS = int(os.environ.get('YPB_TOPK_REPS', '50'))
t0 = time.perf_counter()
for _ in range(REPS):
    out = detect_postprocess(preds, 300, 80)
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / REPS * 1e3
res = {"wall_ms_per_call": round(wall, 4)}
try:
    st = torch.cuda.Stream(dev)
    with torch.cuda.stream(st):
        detect_postprocess(preds, 300, 80)
        st.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, stream=st):
            out2 = detect_postprocess(preds, 300, 80)
        for _ in range(3):
            gr.replay()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(20):
            gr.replay()
        b.record(st)
        st.synchronize()
        res["gpu_ms_per_call_graph"] = round(a.elapsed_time(b) / 20, 4)
        res["graph_equals_eager"] = bool(torch.equal(out, out2))
except Exception as exc:  # noqa: BLE001
    res["graph_error"] = f"{type(exc).__name__}: {exc}"[:300]
print(json.dumps(res))
