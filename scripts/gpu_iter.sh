#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2', d['ms_per_step'], {k:(round(v['ms'],4), round(v['frac'],3)) for k,v in d['roofline_kernels'].items()}, 'e2e', d['e2e']['ms_per_step'])"
timeout 400 python scripts/bench_next.py c2 2>&1 | tail -1 | tee gpurun_out/bench_next.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d.items():
    if 'conv_grid' in k or 'plan' in k: print(k, v)"
