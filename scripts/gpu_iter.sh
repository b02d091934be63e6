#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | grep -E "^E  |passed|failed|Error|^tests" | head -40 | tee gpurun_out/pytest_gpu.log
timeout 400 python scripts/bench_next.py c2 2>&1 | tail -1 | tee gpurun_out/bench_next.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d.items():
    if 'conv_grid' in k or 'from_' in k: print(k, v)"
