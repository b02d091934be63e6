#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_fwd|conv_tc_wgrad" -s 2 -c 3 -o gpurun_out/prof_tc -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out/prof_tc.ncu-rep gpurun_out/launches.csv
