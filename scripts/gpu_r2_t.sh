#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/bench_variants.py c2x256 0,15,16 gpurun_out/r2t_variants_c2x256.json 2>&1 | grep -E "^\{|rror" | cut -c1-200
timeout 600 python scripts/bench_variants.py c2x128 0,13,17 gpurun_out/r2t_variants_c2x128.json 2>&1 | grep -E "^\{|rror" | cut -c1-200
timeout 600 python scripts/bench_variants.py c2 0,17 gpurun_out/r2t_variants_c2.json 2>&1 | grep -E "^\{|rror" | grep -v wgrad | cut -c1-200
