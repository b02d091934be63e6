#!/bin/bash
# Round 2, end of round: tests, smoke and bench lines of the final tree
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2end_pytest_gpu.log
tail -2 gpurun_out/r2end_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2end_smoke.log 2>&1; tail -2 gpurun_out/r2end_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2end_bench.err | tail -1 > gpurun_out/r2end_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2end_bench_reference.json
for c in c1 c2f32 c2x128 c2x256; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2end_bench_$c.json
done
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2end_bench_c5.json
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --graph 2>/dev/null | tail -1 > gpurun_out/r2end_bench_c3_graph.json
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --profile --no-cpu-baseline > gpurun_out/r2end_bench_c3.json 2> gpurun_out/r2end_c3_profile.txt
for f in r2end_bench_c5 r2end_bench_c3_graph r2end_bench r2end_bench_c1 r2end_bench_c2f32 r2end_bench_c2x128 r2end_bench_c2x256; do python - "$f" <<'PY'
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
