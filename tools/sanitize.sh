# compute-sanitizer over the torch-free C-ABI harness (tests/cabi_device_harness.cu): memcheck, racecheck, synccheck on the small
# sizes, in both decode placements (fused into the class scan / separate survivor-decode kernel + persistent TMA scan).
# usage (GPU box): bash tools/sanitize.sh > gpurun_out/r02_sanitizer.txt 2>&1
H=ultralytics_pro_b200/_lib/ypb_cabi_harness
CS=/usr/local/cuda/bin/compute-sanitizer
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
run() { echo "== $*"; "$@" 2>&1 | tail -n ${TAILN:-30}; echo "rc=${PIPESTATUS[0]}"; }
TAILN=40 run timeout 90 $H
TAILN=25 run env YPB_FUSE_DECODE=0 timeout 60 $H --small
run timeout 70 $CS --tool memcheck --error-exitcode 9 --print-limit 20 $H --small
run env YPB_FUSE_DECODE=0 timeout 70 $CS --tool memcheck --error-exitcode 9 --print-limit 20 $H --small
run timeout 90 $CS --tool racecheck --racecheck-report analysis --error-exitcode 9 --print-limit 20 $H --small
run env YPB_FUSE_DECODE=0 timeout 60 $CS --tool synccheck --error-exitcode 9 --print-limit 20 $H --small
