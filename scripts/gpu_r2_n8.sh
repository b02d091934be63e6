#!/bin/bash
# Round 2: the driver's scaling launch at N = 8 (C2 weak line + strong_c4 + train_c3 sub-records), ours and the reference arm
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > gpurun_out/r2_n8_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -3 gpurun_out/r2_bench_n8.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2_bench_n8.json") if l.startswith("{")][-1])
    print("n8 main", d["n_gpus"], round(d["ms_per_step"],3), round(d["value"]/1e9,3), "Gvox/s e2e", (d.get("e2e") or {}).get("ms_per_step"), (d.get("e2e") or {}).get("host_link_probe"))
    for s in ("strong_c4","train_c3"):
        if s in d: print("   ", s, round(d[s]["ms_per_step"],3), round(d[s]["value"]/1e6,1), "Mvox/s", d[s]["config"].get("launch_mode","")[:60], (d[s].get("clocks") or {}).get("reasons"))
except Exception as e:
    print("FAILED", e)
PY
