# N=2 at K=20: which piece of the one-sided gather costs what (YPB_PEER_DEBUG bit mask: 1 local ring only, 2 no fence, 4 no ack wait)
run() { echo "== $* $EXTRA"; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 $EXTRA 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f us/step %.2f verified %s' % (d['value'], d['ms_per_step']*1e3, d['gather_verified_against_nccl']))
    elif 'rror' in l or 'unavailable' in l: print(l.strip()[:300])
"; }
EXTRA="" run YPB_BENCH_QUICK=1
EXTRA="" run YPB_BENCH_QUICK=1 YPB_PEER_DEBUG=1
EXTRA="" run YPB_BENCH_QUICK=1 YPB_PEER_DEBUG=2
EXTRA="" run YPB_BENCH_QUICK=1 YPB_PEER_DEBUG=4
EXTRA="" run YPB_BENCH_QUICK=1 YPB_PEER_DEBUG=7
EXTRA="" run YPB_BENCH_QUICK=1 YPB_PEER_DEBUG=7 YPB_BENCH_NO_CONSUME=1
EXTRA="--gather none" run YPB_BENCH_QUICK=1
