#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "values_and_gradients or identity_map or fused_block or conv_bn_act or c5_ or c3_ or c1_ or fp32 or determinism or tensor_memory or fused_backward or empty" 2>&1 | tail -3
timeout 600 python scripts/time_fused.py 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/r2z_time_narrow.jsonl
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2z_bench_c5.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2z_bench_c5.json").read().strip().splitlines()[-1])
k=d.get("roofline_kernels") or {}
print("c5", round(d["ms_per_step"],3), {n:(round(x["ms"],3), round(x["frac"],3)) for n,x in k.items()}, d.get("roofline_step"))
PY
