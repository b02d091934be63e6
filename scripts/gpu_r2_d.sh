#!/bin/bash
# Round 2, call D: full GPU suite on the new kernel-map kernel and defaults, wgrad shapes, C3 step, default bench, other configs.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2d_pytest_gpu.log
tail -6 gpurun_out/r2d_pytest_gpu.log
timeout 300 python scripts/bench_variants.py c2 0,7,1 gpurun_out/r2d_variants_c2.json 2>&1 | tail -6
timeout 300 python scripts/bench_plan.py c2 2>&1 | tail -1 > gpurun_out/r2d_bench_plan_c2.json; cut -c1-700 gpurun_out/r2d_bench_plan_c2.json
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --profile --no-cpu-baseline > gpurun_out/r2d_bench_c3.json 2> gpurun_out/r2d_c3_profile.txt
grep -E "^\[profile\]" gpurun_out/r2d_c3_profile.txt | head -24 | cut -c1-130
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --graph --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2d_bench_c3_graph.json
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --graph --unfused-bn --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2d_bench_c3_graph_unfused.json
for f in r2d_bench_c3 r2d_bench_c3_graph r2d_bench_c3_graph_unfused; do python -c "
import json,sys
d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],3), d['config']['launch_mode'][:60], d['gpu_launches'])"; done
timeout 900 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2d_bench.json
for c in c1 c2f32 c2x128 c2x256; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2d_bench_$c.json
done
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2d_bench_c5.json
timeout 600 python scripts/bench_plan.py c5 2>&1 | tail -1 > gpurun_out/r2d_bench_plan_c5.json
for f in r2d_bench r2d_bench_c1 r2d_bench_c2f32 r2d_bench_c2x128 r2d_bench_c2x256 r2d_bench_c5; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    k=d.get("roofline_kernels") or {}
    print(f, round(d["ms_per_step"],3), "ms", {n:(round(v["ms"],3), round(v["frac"],3), round(v["compulsory"]["frac"],3)) for n,v in k.items()}, round((d.get("roofline_step") or {}).get("frac",0),3), "plan", round(d["config"]["plan_build_ms"],2), (d.get("clocks") or {}).get("reasons"))
    for s in ("strong_c4","train_c3"):
        if s in d: print("   ", s, round(d[s].get("ms_per_step",0),3), (d[s].get("roofline") or {}).get("frac"), (d[s].get("clocks") or {}).get("reasons"))
except Exception as e:
    print(f, "FAILED", e)
PY
done
