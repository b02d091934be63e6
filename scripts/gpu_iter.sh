#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --config c2x128 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c2x128.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('c2x128', d['ms_per_step'], {k:(round(v['ms'],3), round(v['frac'],3)) for k,v in d['roofline_kernels'].items()}, d['roofline_step']['frac'])"
