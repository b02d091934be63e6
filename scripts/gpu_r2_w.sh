#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "fp32 or c1_ or values_and_gradients or fused_block or conv_bn_act or empty" 2>&1 | tail -3
timeout 600 python bench.py --config c1 --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2w_bench_c1.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2w_bench_c1.json").read().strip().splitlines()[-1])
k=d.get("roofline_kernels") or {}
print("c1", round(d["ms_per_step"],3), {n:(round(v["ms"],3), round(v["frac"],3)) for n,v in k.items()}, d.get("roofline_step"))
PY
