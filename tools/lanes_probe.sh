run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 400 --warmup 10 --no-extra --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('value %.0f us/step %.2f single %.1f scan %.1f' % (d['value'], d['ms_per_step']*1e3, r['single_stream_step_ms']*1e3, r['launch_ms']*1e3))
    elif 'rror' in l: print(l.strip()[:200])
"; }
run YPB_SCAN_TMA=0
run YPB_SCAN_TMA=1
run YPB_TMA_SMEM_KB=100
run YPB_TMA_SMEM_KB=60 YPB_TMA_NCW=8
run YPB_TMA_SMEM_KB=100 YPB_TMA_LB=16
