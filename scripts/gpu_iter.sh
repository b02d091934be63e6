#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"kmap_build_kernel|csr_fill|tile_mask" -s 2 -c 3 -o gpurun_out/prof_kmap -f python scripts/bench_plan.py c2 > gpurun_out/kmap_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"bn_stats_partial|bn_apply|bn_backward|pool_rows|gather_rows" -s 8 -c 8 -o gpurun_out/prof_norm -f python scripts/bench_next.py c2 > gpurun_out/norm_under_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
