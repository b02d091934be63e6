#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_c3.json | cut -c1-900
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3.csv python bench.py --config c3 --steps 1 --warmup 1 > gpurun_out/c3_under_ncu.log 2>&1
tail -1 gpurun_out/c3_under_ncu.log | cut -c1-200
