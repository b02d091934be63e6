#!/bin/bash
# Round 2, call C (N GPUs of one box): by-grid equivalence on two devices and over NCCL, the default bench line at N ranks
# (C2 weak + strong_c4 + train_c3 as ONE CUDA graph with the NCCL all-reduces captured), SyncBatchNorm check.
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > gpurun_out/r2c_gpu.csv 2>&1
timeout 600 python -m pytest tests/test_gpu_at_size.py -m gpu -q -k "by_grid or c1_fp32" 2>&1 | tail -5 | tee gpurun_out/r2c_pytest_two_devices.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 scripts/check_by_grid_nccl.py 2>&1 | tail -3 | tee gpurun_out/r2c_by_grid_nccl_n$N.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 scripts/check_syncbn.py 2>&1 | tail -2 | cut -c1-300 | tee gpurun_out/r2c_syncbn_n$N.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/r2c_bench_n$N.json 2> gpurun_out/r2c_bench_n$N.err
tail -4 gpurun_out/r2c_bench_n$N.err
python - "$N" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2c_bench_n{n}.json").read().strip().splitlines()[-1])
    print("C2 weak", round(d["ms_per_step"], 3), "ms", round(d["value"] / 1e9, 3), "Gvox/s; e2e", round(d["e2e"]["ms_per_step"], 2), "ms", d["e2e"]["host_link_probe"])
    for k in ("strong_c4", "train_c3"):
        s = d.get(k, {})
        print(k, {kk: s.get(kk) for kk in ("ms_per_step", "value", "error")}, (s.get("config") or {}).get("launch_mode"), (s.get("clocks") or {}))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29554 bench.py --gpus $N --config c3 --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2c_bench_c3_eager_n$N.json
python -c "
import json; d=json.loads(open('gpurun_out/r2c_bench_c3_eager_n$N.json').read().strip().splitlines()[-1]); print('c3 eager', d['ms_per_step'], d['config']['launch_mode'])"
