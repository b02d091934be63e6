#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 0 5 6; do
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --wgrad-variant $v 2>/dev/null | tail -1 > gpurun_out/r2y_bench_c4_wv$v.json
python - $v <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads(open(f"gpurun_out/r2y_bench_c4_wv{v}.json").read().strip().splitlines()[-1])
k=d.get("roofline_kernels") or {}
print("c4 wgrad variant", v, round(d["ms_per_step"],3), {n:(round(x["ms"],3), round(x["frac"],3)) for n,x in k.items()}, (d.get("clocks") or {}))
PY
done
for v in 0 5; do
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline --wgrad-variant $v 2>/dev/null | tail -1 > gpurun_out/r2y_bench_c5_wv$v.json
python - $v <<'PY'
import json,sys
v=sys.argv[1]
d=json.loads(open(f"gpurun_out/r2y_bench_c5_wv{v}.json").read().strip().splitlines()[-1])
k=d.get("roofline_kernels") or {}
print("c5 wgrad variant", v, round(d["ms_per_step"],3), {n:(round(x["ms"],3), round(x["frac"],3)) for n,x in k.items()})
PY
done
