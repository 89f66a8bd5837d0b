# N=2 weak-scaling step time against the number of lanes (diagnostic; run under gpurun --gpus 2)
run() { echo "== $* $EXTRA"; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 600 --warmup 20 $EXTRA 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f us/step %.2f verified %s' % (d['value'], d['ms_per_step']*1e3, d['gather_verified_against_nccl']))
    elif 'rror' in l: print(l.strip()[:200])
"; }
EXTRA="--lanes 5" run YPB_BENCH_QUICK=1
EXTRA="--lanes 7" run YPB_BENCH_QUICK=1
EXTRA="--lanes 10" run YPB_BENCH_QUICK=1
EXTRA="--lanes 7 --gather none" run YPB_BENCH_QUICK=1
EXTRA="--lanes 5 --gather-lag 0" run YPB_BENCH_QUICK=1
