#!/bin/bash
# Round 2, call B: full parity suite, more pipeline shapes, C3 step profile (default vs round-1 kernels), plan-build pieces.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2b_pytest_parity.log
tail -6 gpurun_out/r2b_pytest_parity.log
timeout 300 python scripts/bench_variants.py c2 1,0,7,8,9,10 gpurun_out/r2b_variants_c2.json 2>&1 | tail -8
timeout 300 python scripts/bench_variants.py c2x128 1,0,7,8,9,10 gpurun_out/r2b_variants_c2x128.json 2>&1 | tail -8
timeout 300 python scripts/bench_variants.py c2x256 1,0 gpurun_out/r2b_variants_c2x256.json 2>&1 | tail -3
timeout 300 python scripts/bench_plan.py c2 2>&1 | tail -1 > gpurun_out/r2b_bench_plan_c2.json; cat gpurun_out/r2b_bench_plan_c2.json
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --profile --no-cpu-baseline > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_c3_profile.txt
grep -E "^\[profile\]" gpurun_out/r2b_c3_profile.txt | head -40
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --profile --no-cpu-baseline --variant 1 > gpurun_out/r2b_bench_c3_v1.json 2> gpurun_out/r2b_c3_profile_v1.txt
grep -E "^\[profile\]" gpurun_out/r2b_c3_profile_v1.txt | head -14
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --graph --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2b_bench_c3_graph.json
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --graph --unfused-bn --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2b_bench_c3_graph_unfused.json
for f in r2b_bench_c3 r2b_bench_c3_v1 r2b_bench_c3_graph r2b_bench_c3_graph_unfused; do python -c "
import json,sys
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],3), d['config']['launch_mode'][:60], d['gpu_launches'])"; done
