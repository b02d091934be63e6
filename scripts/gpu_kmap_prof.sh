#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pred_gather" 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"kmap_build" -s 3 -c 1 -o gpurun_out/prof_kmap -f python scripts/bench_plan.py c2 > gpurun_out/kmap_under_ncu.log 2>&1
tail -2 gpurun_out/kmap_under_ncu.log | cut -c1-400
