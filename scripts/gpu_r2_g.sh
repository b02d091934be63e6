#!/bin/bash
set -u
mkdir -p gpurun_out
for cfg in "16 16 5 3000" "32 32 3 3000" "16 16 3 40000" "16 32 3 3000"; do
  echo "== $cfg"
  FVC_DEBUG_WAITS=1 timeout 120 python scripts/dbg_fused.py $cfg 2>&1 | grep -E "fvc debug|rel err|rows" | cut -c1-300
done
bash scripts/gpu_r2_f.sh
FVC_FUSED_BACKWARD=0 timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2f_bench_c5_unfused.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f_bench_c5_unfused.json").read().strip().splitlines()[-1])
print("c5 unfused", round(d["ms_per_step"],3), d.get("phase_ms"))
PY
