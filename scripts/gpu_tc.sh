#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "identity_map" 2>&1 | tail -30 | tee gpurun_out/pytest_tc_identity.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "values_and_gradients or determinism or forced or nn_modules" 2>&1 | tail -40 | tee gpurun_out/pytest_tc_values.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_tc.log
