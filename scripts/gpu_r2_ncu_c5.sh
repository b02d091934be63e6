#!/bin/bash
# Round 2: ncu --set full of the narrow-layer kernels on one grid of C5 (forward, fused backward) and of the fp32 kernels on C1
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:"conv_tc_fwd|conv_tc_bwd_fused" -s 4 -c 2 -o gpurun_out/r2_prof_c5 -f python bench.py --config c5 --grids 1 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_under_ncu_c5.log 2>&1
ncu -i gpurun_out/r2_prof_c5.ncu-rep --page raw --csv > gpurun_out/r2_prof_c5_raw.csv 2>/dev/null; rm -f gpurun_out/r2_prof_c5.ncu-rep
timeout 900 ncu --set full --clock-control none -k regex:"conv_tc_fwd|conv_tc_wgrad" -s 6 -c 3 -o gpurun_out/r2_prof_c2f32 -f python bench.py --config c2f32 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_under_ncu_c2f32.log 2>&1
ncu -i gpurun_out/r2_prof_c2f32.ncu-rep --page raw --csv > gpurun_out/r2_prof_c2f32_raw.csv 2>/dev/null; rm -f gpurun_out/r2_prof_c2f32.ncu-rep
ls -la gpurun_out | grep "r2_prof"
