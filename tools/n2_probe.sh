run() { echo "== $* $EXTRA"; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 400 --warmup 10 $EXTRA 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f us/step %.2f verified %s strong %s' % (d['value'], d['ms_per_step']*1e3, d['gather_verified_against_nccl'], json.dumps(d.get('strong_scaling'))[60:160]))
    elif 'rror' in l: print(l.strip()[:200])
"; }
timeout 200 python -m pytest tests/test_multi_gpu.py -x -q 2>&1 | tail -3
EXTRA="" run A=1
EXTRA="" run YPB_PEER_PUSH_SPLIT=1
