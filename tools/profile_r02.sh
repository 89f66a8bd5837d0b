#!/bin/bash
# Round-2 profiling recipe (run under gpurun, ONE GPU).  Numbers printed by runs under ncu are never bench values.
set -x
K='regex:scan_classes|decode_tiles|sort_suppress|fast_nms_cluster|decode_dense|filter_from_dense|compact_results'
# 1. launch list of the bench command (cold-cache, serialised: compare SHARES)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 500 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02_launches_bench.log 2>&1
# 2. full capture of the three kernels of the step (fp32 C2, B=64), one lane so the instances are the B=64 ones
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:scan_classes_kernel|decode_tiles|sort_suppress' -s 30 -c 6 \
  -o gpurun_out/r02_step_f32 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra --lanes 1 > gpurun_out/r02_step_f32.log 2>&1
# 3. the same for the bf16 head
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:scan_classes_kernel|decode_tiles' -s 30 -c 4 \
  -o gpurun_out/r02_step_bf16 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra --lanes 1 --dtype bf16 > gpurun_out/r02_step_bf16.log 2>&1
# 4. the persistent TMA scan (fp32 and bf16) and the cluster Fast-NMS kernel (C5)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_classes_tma -s 12 -c 2 -o gpurun_out/r02_scan_tma_f32 \
  env YPB_SCAN_TMA=1 python tools/scan_tune.py c2_v8x_640_b64 f32 > gpurun_out/r02_scan_tma_f32.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_classes_tma -s 12 -c 2 -o gpurun_out/r02_scan_tma_bf16 \
  env YPB_SCAN_TMA=1 python tools/scan_tune.py c2_v8x_640_b64 bf16 > gpurun_out/r02_scan_tma_bf16.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fast_nms_cluster -s 6 -c 2 -o gpurun_out/r02_fast_nms_c5 \
  python tools/scan_tune.py c5_obb_1024_b16 f32 > gpurun_out/r02_fast_nms_c5.log 2>&1
# 5. dense decode, 16-bit
timeout 300 ncu --set full --clock-control none --import-source on -k regex:decode_dense -s 3 -c 2 -o gpurun_out/r02_dense_bf16 \
  python tools/dense16_probe.py > gpurun_out/r02_dense_bf16.log 2>&1
ls -la gpurun_out/*.ncu-rep
# 6. gpurun brings back at most 64 MiB: extract the raw pages on the box, keep only the smaller reports
for r in r02_step_f32 r02_step_bf16 r02_scan_tma_f32 r02_scan_tma_bf16 r02_fast_nms_c5 r02_dense_bf16; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/r02_scan_tma_bf16.ncu-rep --page source --csv > gpurun_out/r02_scan_tma_bf16_source.csv 2>/dev/null
ncu -i gpurun_out/r02_fast_nms_c5.ncu-rep --page source --csv > gpurun_out/r02_fast_nms_c5_source.csv 2>/dev/null
rm -f gpurun_out/r02_step_bf16.ncu-rep gpurun_out/r02_dense_bf16.ncu-rep gpurun_out/r02_step_f32.ncu-rep
du -sh gpurun_out
