#!/bin/bash
# Round 2: the driver's scaling launch at N = 2 and N = 4 (final tree), plus the 2-device tests
set -u
mkdir -p gpurun_out
for n in 2 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/r2_bench_n{n}.json") if l.startswith("{")][-1])
    print("n", n, "main", round(d["ms_per_step"],3), round(d["value"]/1e9,3), "Gvox/s e2e", (d.get("e2e") or {}).get("ms_per_step"))
    for s in ("strong_c4","train_c3"):
        if s in d: print("   ", s, round(d[s]["ms_per_step"],3), round(d[s]["value"]/1e6,1), "Mvox/s", d[s]["config"].get("launch_mode","")[:50], (d[s].get("clocks") or {}).get("reasons"))
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/r2_bench_n{n}.err").read()[-1500:])
PY
done
timeout 600 python -m pytest tests/test_gpu_at_size.py -m gpu -q -k "two_devices" 2>&1 | tail -2 | tee gpurun_out/r2_pytest_two_devices.log
