#!/bin/bash
# Diagnostic: rebuild the library with different streaming-load flavours on the GPU box and report kernel times.
set -e
for mode in 0 1 2 3 4; do
  YPB_EXTRA_NVCC="-DYPB_LOAD_MODE=$mode" python -m ultralytics_pro_b200.build --force > /dev/null 2>/tmp/build.err || { tail -5 /tmp/build.err; continue; }
  python bench.py --steps 600 --warmup 20 --no-cpu-baseline --lanes 1 > /tmp/b.json 2>/tmp/b.err || { tail -3 /tmp/b.err; continue; }
  python - "$mode" <<PY
import json,sys
d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1])
print("load mode", sys.argv[1], "value", round(d["value"]), "scan_ms", round(d["roofline"]["launch_ms"],4), "frac", round(d["roofline"]["frac"],3), "dense", round(d["decode_dense"]["launch_ms"],4), round(d["decode_dense"]["frac"],3))
PY
done
python -m ultralytics_pro_b200.build --force > /dev/null
