#!/bin/bash
# Round-end single-GPU evidence run: all GPU tests, smoke, both bench arms, every named configuration, ncu launch list,
# ncu --set full captures of the hot kernels (convolution, kernel-map build).  Everything is wrapped in `timeout`.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.json
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench.json
cut -c1-400 gpurun_out/bench.json
for c in c1 c2f32 c2x128 c3; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$c.json
done
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --graph 2>&1 | tail -1 > gpurun_out/bench_c3_graph.json
timeout 900 python bench.py --config c5 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c5.json
timeout 900 python bench.py --config c5f32 --grids 2 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c5f32_2grids.json
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_c4.json
timeout 300 python scripts/bench_next.py c2 2>&1 | tail -1 > gpurun_out/bench_next.json
timeout 300 python scripts/bench_plan.py c2 2>&1 | tail -1 > gpurun_out/bench_plan_c2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_fwd|conv_tc_wgrad" -s 2 -c 3 -o gpurun_out/prof_tc -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"kmap_build|dilate_leaves|bn_stats_partial|bn_apply" -c 6 -o gpurun_out/prof_aux -f python scripts/bench_next.py c2 > gpurun_out/aux_under_ncu.log 2>&1
for f in bench_c1 bench_c2f32 bench_c2x128 bench_c3 bench_c3_graph bench_c5 bench_c5f32_2grids bench_c4; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    k=d.get("roofline_kernels") or {}
    print(f, round(d["ms_per_step"],3), "ms", {n:(round(v["ms"],3), round(v["frac"],3)) for n,v in k.items()}, (d.get("roofline_step") or {}).get("frac"))
except Exception as e:
    print(f, "FAILED", e)
PY
done
ls gpurun_out
