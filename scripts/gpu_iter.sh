#!/bin/bash
# Iteration check: parity tests, default bench, fp32 / small-channel / 128-channel configs.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c2.json | cut -c1-1500
timeout 600 python bench.py --config c1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c1.json | cut -c1-2600
for v in 1 2 3; do
FVC_TC_VARIANT=$v timeout 900 python bench.py --config c5 --grids 1 --steps 3 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c5_v$v.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c5 variant', {k:v['ms'] for k,v in d['roofline_kernels'].items()})"
done
