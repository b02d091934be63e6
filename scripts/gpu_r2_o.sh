#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "identity_map or values_and_gradients or fused_block or conv_bn_act" 2>&1 | tail -2
timeout 600 python scripts/time_fused.py 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/r2o_time_ts.jsonl
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2o_bench_c5.err | tail -1 > gpurun_out/r2o_bench_c5.json
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2o_bench_c5.json").read().strip().splitlines()[-1])
    k=d.get("roofline_kernels") or {}
    print("c5", round(d["ms_per_step"],3), {n:(round(v["ms"],3), round(v["frac"],3)) for n,v in k.items()}, d.get("roofline_step"), d.get("phase_ms"))
except Exception as e:
    print("c5 FAILED", e); print(open("gpurun_out/r2o_bench_c5.err").read()[-1500:])
PY
