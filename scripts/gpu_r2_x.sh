#!/bin/bash
set -u
mkdir -p gpurun_out
for c in c2 c2x128 c2x256; do
  timeout 600 python scripts/bench_variants.py $c 0 gpurun_out/r2x_variants_$c.json 2>&1 | grep -E "wgrad|rror" | cut -c1-220
done
timeout 900 python -m pytest tests -m gpu -q -x -k "values_and_gradients or at_size or c2_full or c4_ or c3_ or c1_ or fp32 or nn_modules or empty or host_pipelined" 2>&1 | tail -3
