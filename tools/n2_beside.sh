# N=2: consumer forked beside the next step (default) against the consumer behind the suppression kernel (diagnostic; gpurun --gpus 2)
run() { echo "== $* $EXTRA"; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $EXTRA 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f us/step %.2f verified %s' % (d['value'], d['ms_per_step']*1e3, d['gather_verified_against_nccl']))
    elif 'rror' in l or 'unavailable' in l: print(l.strip()[:300])
"; }
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -n 5
EXTRA="--steps 20 --warmup 5" run YPB_BENCH_QUICK=1
EXTRA="--steps 20 --warmup 5" run YPB_BENCH_QUICK=1 YPB_BENCH_TAIL_CONSUME=1
EXTRA="--steps 20 --warmup 5" run YPB_BENCH_QUICK=1
EXTRA="--steps 600 --warmup 20" run YPB_BENCH_QUICK=1
EXTRA="--steps 600 --warmup 20" run YPB_BENCH_QUICK=1 YPB_BENCH_TAIL_CONSUME=1
EXTRA="--steps 600 --warmup 20 --gather none" run YPB_BENCH_QUICK=1
