#!/bin/bash
# Round 2, call A: parity tests of the new kernels (small cases first), pipeline-shape variants on C2 / C2x128, the default bench line.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/r2a_gpu.csv 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --deselect tests/test_gpu_parity.py::test_full_size_properties_on_the_bench_workload 2>&1 | tail -40 > gpurun_out/r2a_pytest_parity.log
tail -5 gpurun_out/r2a_pytest_parity.log
timeout 300 python scripts/bench_variants.py c2 1,0,2,3,4,5,6 gpurun_out/r2a_variants_c2.json 2>&1 | tail -8
timeout 300 python scripts/bench_variants.py c2x128 1,0,2,3,4,5,6 gpurun_out/r2a_variants_c2x128.json 2>&1 | tail -8
timeout 1200 python -m pytest tests/test_gpu_at_size.py tests/test_gpu_parity.py::test_full_size_properties_on_the_bench_workload -m gpu -q --durations=10 2>&1 | tail -60 > gpurun_out/r2a_pytest_at_size.log
tail -25 gpurun_out/r2a_pytest_at_size.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2a_smoke.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) 2> gpurun_out/r2a_bench.err | tail -1 > gpurun_out/r2a_bench.json
tail -5 gpurun_out/r2a_bench.err
cut -c1-600 gpurun_out/r2a_bench.json
