#!/bin/bash
# Round 2, final call: full validation of the tree -- GPU tests, default bench line, launch list, ncu captures of the hot kernels of
# the 64 / 128 / 256-channel shapes (exported to CSV on the box; the .ncu-rep files are deleted: gpurun_out/ is capped), other configs.
set -u
mkdir -p gpurun_out
export_rep () {  # $1 = report stem
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass > gpurun_out/$1_source.csv 2>/dev/null
  gzip -f gpurun_out/$1_source.csv
  rm -f gpurun_out/$1.ncu-rep
}
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2fin_pytest_gpu.log
tail -5 gpurun_out/r2fin_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2fin_bench.err | tail -1 > gpurun_out/r2fin_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2fin_bench_reference.json
for c in c1 c2f32 c2x128 c2x256; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2fin_bench_$c.json
done
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2fin_bench_c5.json
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --graph --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r2fin_bench_c3_graph.json
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --profile --no-cpu-baseline > gpurun_out/r2fin_bench_c3.json 2> gpurun_out/r2fin_c3_profile.txt
for f in r2fin_bench r2fin_bench_c1 r2fin_bench_c2f32 r2fin_bench_c2x128 r2fin_bench_c2x256 r2fin_bench_c5 r2fin_bench_c3_graph; do python - "$f" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    k=d.get("roofline_kernels") or {}
    print(f, round(d["ms_per_step"],3), "ms", {n:(round(v["ms"],3), round(v["frac"],3), round(v["compulsory"]["frac"],3)) for n,v in k.items()}, round((d.get("roofline_step") or {}).get("frac",0),3), (d.get("clocks") or {}).get("reasons"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
    for s in ("strong_c4","train_c3"):
        if s in d: print("   ", s, round(d[s].get("ms_per_step",0),3), (d[s].get("roofline") or {}).get("frac"), (d[s].get("clocks") or {}).get("reasons"))
except Exception as e:
    print(f, "FAILED", e)
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_fwd|conv_tc_wgrad" -s 6 -c 3 -o gpurun_out/r2fin_prof_c2 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sub-records > gpurun_out/r2fin_under_ncu_c2.log 2>&1
export_rep r2fin_prof_c2
for c in c2x128 c2x256; do
  timeout 900 ncu --set full --clock-control none -k regex:"conv_tc_fwd|conv_tc_wgrad" -s 6 -c 3 -o gpurun_out/r2fin_prof_$c -f python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2fin_under_ncu_$c.log 2>&1
  ncu -i gpurun_out/r2fin_prof_$c.ncu-rep --page raw --csv > gpurun_out/r2fin_prof_${c}_raw.csv 2>/dev/null
  rm -f gpurun_out/r2fin_prof_$c.ncu-rep
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2fin_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2fin_under_ncu_launches.log 2>&1
du -sh gpurun_out
