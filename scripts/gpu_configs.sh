#!/bin/bash
set -u
mkdir -p gpurun_out
for c in c1 c2x128 c3; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$c.json | cut -c1-2500
done
timeout 900 python bench.py --config c5 --grids 2 --steps 3 --warmup 1 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_c5_2grids.json | cut -c1-2500
