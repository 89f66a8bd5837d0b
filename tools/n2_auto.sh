# N=2 (gpurun --gpus 2): the 2-GPU tests, then the bench in quick mode with the consumer placement picked by trial (auto) and fixed (beside)
run() { echo "== $* $EXTRA"; env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $EXTRA 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f us/step %.2f verified %s pick %s' % (d['value'], d['ms_per_step']*1e3, d['gather_verified_against_nccl'], d.get('gather_consumer')))
    elif 'rror' in l or 'unavailable' in l: print(l.strip()[:300])
"; }
timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -n 3
EXTRA="--steps 20 --warmup 5" run YPB_BENCH_QUICK=1 YPB_BENCH_CONSUMER=auto
EXTRA="--steps 20 --warmup 5" run YPB_BENCH_QUICK=1
