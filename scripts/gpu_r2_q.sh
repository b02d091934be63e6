#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/bench_variants.py c2 0,11 gpurun_out/r2q_variants_c2.json 2>&1 | grep -E "^\{|rror" | cut -c1-200
timeout 600 python scripts/bench_variants.py c2x128 0,11 gpurun_out/r2q_variants_c2x128.json 2>&1 | grep -E "^\{|rror" | cut -c1-200
