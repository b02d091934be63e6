#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error" | head -20 | tee gpurun_out/pytest_gpu.log
for c in c2 c3; do
timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$c', d['ms_per_step'], {k:(round(v['ms'],4), round(v['frac'],3)) for k,v in (d.get('roofline_kernels') or {}).items()}, (d.get('roofline_step') or {}).get('frac'))"
done
