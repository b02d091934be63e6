#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c2.json | cut -c1-300
for c in c1 c2f32; do
timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$c.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$c', d['ms_per_step'], {k:(round(v['ms'],4), round(v['frac'],3)) for k,v in d['roofline_kernels'].items()}, d['roofline_step'])"
done
