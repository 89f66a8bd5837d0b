for dt in f32 bf16; do
  YPB_SCAN_TMA=0 timeout 100 python tools/scan_tune.py c2_v8x_640_b64 $dt
  for lb in 16 8; do for ncw in 8 12 16; do
    YPB_TMA_LB=$lb YPB_TMA_NCW=$ncw timeout 100 python tools/scan_tune.py c2_v8x_640_b64 $dt
  done; done
done
