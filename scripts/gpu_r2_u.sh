#!/bin/bash
set -u
mkdir -p gpurun_out
for c in c2 c2x128 c2x256; do
  timeout 600 python scripts/bench_variants.py $c 0 gpurun_out/r2u_variants_$c.json 2>&1 | grep -E "^\{|rror" | cut -c1-220
done
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "values_and_gradients or identity_map or pipeline_variants or fused_block" 2>&1 | tail -3
