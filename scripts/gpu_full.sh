#!/bin/bash
# Full single-GPU round check: all GPU tests, smoke, both bench arms, ncu launch list + full capture.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_fwd|conv_tc_wgrad" -s 2 -c 3 -o gpurun_out/prof_tc -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1
ls gpurun_out
