#!/bin/bash
# Round 2, call E: ncu evidence.  Full-set captures of the hot kernels of the default bench step (forward, weight gradient, dgrad),
# of a one-CTA-per-SM forward shape (why are few big CTAs slow?), of the kernel-map kernel; the launch list of the bench command.
# The .ncu-rep files are exported to CSV on the box (raw metrics + per-instruction source page) and deleted: gpurun_out/ is capped at 64 MiB.
set -u
mkdir -p gpurun_out
export_rep () {  # $1 = report stem
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass > gpurun_out/$1_source.csv 2>/dev/null
  gzip -f gpurun_out/$1_source.csv
  rm -f gpurun_out/$1.ncu-rep
}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_fwd|conv_tc_wgrad" -s 6 -c 3 -o gpurun_out/r2e_prof_c2 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-records > gpurun_out/r2e_under_ncu_c2.log 2>&1
export_rep r2e_prof_c2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_fwd" -s 4 -c 1 -o gpurun_out/r2e_prof_c2_v3 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-records --variant 3 > gpurun_out/r2e_under_ncu_v3.log 2>&1
export_rep r2e_prof_c2_v3
timeout 600 ncu --set full --clock-control none -k regex:"kmap_build" -s 2 -c 1 -o gpurun_out/r2e_prof_kmap -f python scripts/bench_plan.py c2 > gpurun_out/r2e_under_ncu_kmap.log 2>&1
export_rep r2e_prof_kmap
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_under_ncu_launches.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "grid_build or from_points or generated_grids" 2>&1 | tail -3
timeout 300 python scripts/bench_plan.py c2 2>&1 | tail -1 > gpurun_out/r2e_bench_plan_c2.json; cut -c1-400 gpurun_out/r2e_bench_plan_c2.json
du -sh gpurun_out; ls -la gpurun_out
