#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_ts" -s 2 -c 1 -o gpurun_out/r2j_prof_ts -f python scripts/prof_ts.py 16 16 5 > gpurun_out/r2j_under_ncu.log 2>&1
tail -3 gpurun_out/r2j_under_ncu.log
ncu -i gpurun_out/r2j_prof_ts.ncu-rep --page raw --csv > gpurun_out/r2j_prof_ts_raw.csv 2>/dev/null
ncu -i gpurun_out/r2j_prof_ts.ncu-rep --page source --csv --print-source sass > gpurun_out/r2j_prof_ts_source.csv 2>/dev/null
gzip -f gpurun_out/r2j_prof_ts_source.csv
rm -f gpurun_out/r2j_prof_ts.ncu-rep
ls -la gpurun_out | grep r2j
