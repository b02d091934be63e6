#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20 | tee gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 300 python bench.py 2>&1 | tail -1 > gpurun_out/bench.json
python -c "
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); print('c2', d['ms_per_step'], d['value'], {k:(round(v['ms'],4), round(v['frac'],3)) for k,v in d['roofline_kernels'].items()}, d['roofline_step']['frac'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])"
