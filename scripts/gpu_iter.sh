#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --config c5 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c5.json
python -c "
import json
d=json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1]); print('c5', d['ms_per_step'], {k:(round(v['ms'],3), round(v['frac'],3)) for k,v in d['roofline_kernels'].items()}, d['roofline_step']['frac'])"
