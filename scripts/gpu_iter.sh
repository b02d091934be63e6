#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20 | tee gpurun_out/pytest_gpu.log
timeout 300 python scripts/bench_plan.py c2 2>&1 | tail -1 | tee gpurun_out/bench_plan_c2.json
timeout 300 python scripts/bench_plan.py c5 2>&1 | tail -1 | tee gpurun_out/bench_plan_c5.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"kmap_build" -s 3 -c 1 -o gpurun_out/prof_kmap -f python scripts/bench_plan.py c2 > gpurun_out/kmap_under_ncu.log 2>&1
ls gpurun_out | head -30
