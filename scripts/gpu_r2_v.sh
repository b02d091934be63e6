#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for c in c2x128 c2x256; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2v_bench_$c.json
done
timeout 900 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2v_bench.json
for f in r2v_bench r2v_bench_c2x128 r2v_bench_c2x256; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    k=d.get("roofline_kernels") or {}
    print(f, round(d["ms_per_step"],3), "ms", {n:(round(v["ms"],3), round(v["frac"],3), round(v["compulsory"]["frac"],3)) for n,v in k.items()}, round((d.get("roofline_step") or {}).get("frac",0),3), (d.get("clocks") or {}).get("reasons"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "traffic", (d.get("roofline") or {}).get("traffic"))
    for s in ("strong_c4","train_c3"):
        if s in d: print("   ", s, round(d[s].get("ms_per_step",0),3), (d[s].get("roofline") or {}).get("frac"), (d[s].get("clocks") or {}).get("reasons"), (d[s].get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print(f, "FAILED", e)
PY
done
