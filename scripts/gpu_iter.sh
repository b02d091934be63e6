#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "batch_norm or column_sums or simple_unet or nn_" 2>&1 | grep -E "^E  |passed|failed|Error" | head -10
timeout 300 python scripts/bench_next.py c2 2>&1 | tail -1 | tee gpurun_out/bench_next.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d.items():
    if 'BatchNorm+ReLU' in k: print(k, v)"
timeout 300 python bench.py --config c3 --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c3', d['ms_per_step'])"
