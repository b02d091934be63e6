#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pipeline_variants" 2>&1 | tail -4
timeout 600 python scripts/bench_variants.py c2 0,13,14 gpurun_out/r2s_variants_c2.json 2>&1 | grep -E "^\{|rror" | grep -v wgrad | cut -c1-200
timeout 600 python scripts/bench_variants.py c2x128 0,13,14 gpurun_out/r2s_variants_c2x128.json 2>&1 | grep -E "^\{|rror" | cut -c1-200
