#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/bench_variants.py c2f32 0,18 gpurun_out/r2aa_variants_c2f32.json 2>&1 | grep -E "^\{|rror" | cut -c1-220
