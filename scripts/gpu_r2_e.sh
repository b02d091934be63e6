#!/bin/bash
# Round 2, call E: ncu evidence.  Full-set captures of the hot kernels of the default bench step (forward, weight gradient, dgrad),
# of a one-CTA-per-SM forward shape (why are few big CTAs slow?), of the kernel-map kernel; the launch list of the bench command.
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_fwd|conv_tc_wgrad" -s 6 -c 3 -o gpurun_out/r2e_prof_c2 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-records > gpurun_out/r2e_under_ncu_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_fwd" -s 4 -c 1 -o gpurun_out/r2e_prof_c2_v3 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-records --variant 3 > gpurun_out/r2e_under_ncu_v3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"kmap_build" -s 2 -c 1 -o gpurun_out/r2e_prof_kmap -f python scripts/bench_plan.py c2 > gpurun_out/r2e_under_ncu_kmap.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2e_under_ncu_launches.log 2>&1
timeout 300 python scripts/bench_variants.py c2 0 gpurun_out/r2e_variants_c2.json 2>&1 | tail -5
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --graph --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2e_bench_c3_graph.json
python -c "
import json; d=json.loads(open('gpurun_out/r2e_bench_c3_graph.json').read().strip().splitlines()[-1]); print('c3 graph', d['ms_per_step'])"
ls -la gpurun_out/*.ncu-rep
