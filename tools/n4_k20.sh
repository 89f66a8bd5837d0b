run() { echo "== $*"; env "$@" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 2>&1 | grep -E "^\{" | cut -c1-120; }
run YPB_BENCH_QUICK=1
run YPB_BENCH_QUICK=1 YPB_BENCH_TAIL_CONSUME=1
run YPB_BENCH_QUICK=1 YPB_BENCH_NO_DEVICE_BARRIER=1
