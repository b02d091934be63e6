#!/bin/bash
# Round 2, call F: the fused backward kernel (narrow layers) -- parity, then C5 with and without it.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_backward" 2>&1 | tail -15 > gpurun_out/r2f_pytest_fused.log
tail -8 gpurun_out/r2f_pytest_fused.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "values_and_gradients or nn_modules or simple_unet or conv_bn_act" 2>&1 | tail -8 > gpurun_out/r2f_pytest_values.log
tail -4 gpurun_out/r2f_pytest_values.log
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2f_bench_c5.err | tail -1 > gpurun_out/r2f_bench_c5.json
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2f_bench_c5.json").read().strip().splitlines()[-1])
    k=d.get("roofline_kernels") or {}
    print("c5", round(d["ms_per_step"],3), {n:(round(v["ms"],3), round(v["frac"],3)) for n,v in k.items()}, d.get("roofline_step"), d.get("phase_ms"))
except Exception as e:
    print("c5 FAILED", e); print(open("gpurun_out/r2f_bench_c5.err").read()[-2000:])
PY
