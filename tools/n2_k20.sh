# N=2 at the driver's K=20 / W=5: with and without the device-side start barrier (diagnostic; run under gpurun --gpus 2)
run() { echo "== $* $EXTRA"; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 $EXTRA 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.0f us/step %.2f verified %s' % (d['value'], d['ms_per_step']*1e3, d['gather_verified_against_nccl']))
    elif 'rror' in l or 'unavailable' in l: print(l.strip()[:200])
"; }
for i in 1 2 3; do
EXTRA="" run YPB_BENCH_QUICK=1
EXTRA="" run YPB_BENCH_QUICK=1 YPB_BENCH_NO_DEVICE_BARRIER=1
done
EXTRA="--gather none" run YPB_BENCH_QUICK=1
EXTRA="--gather none" run YPB_BENCH_QUICK=1 YPB_BENCH_NO_DEVICE_BARRIER=1
