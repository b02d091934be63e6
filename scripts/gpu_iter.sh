#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "batch_norm or column_sums or from_points" 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --profile 2>&1 | tail -30 | cut -c1-200 | tee gpurun_out/bench_c3_profile.txt
