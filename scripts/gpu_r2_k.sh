#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "identity_map or values_and_gradients or fused_block or conv_bn_act" 2>&1 | tail -2
timeout 600 python scripts/time_fused.py 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/r2n_time_ts.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_ts" -s 2 -c 1 -o gpurun_out/r2n_prof_ts -f python scripts/prof_ts.py 16 16 5 > gpurun_out/r2n_under_ncu.log 2>&1
ncu -i gpurun_out/r2n_prof_ts.ncu-rep --page raw --csv > gpurun_out/r2n_prof_ts_raw.csv 2>/dev/null
ncu -i gpurun_out/r2n_prof_ts.ncu-rep --page source --csv --print-source sass > gpurun_out/r2n_prof_ts_source.csv 2>/dev/null
gzip -f gpurun_out/r2n_prof_ts_source.csv
rm -f gpurun_out/r2n_prof_ts.ncu-rep
