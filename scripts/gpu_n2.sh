#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "empty_and_degenerate or full_size" 2>&1 | grep -E "^E |passed|failed|Error" | head -30 | tee gpurun_out/pytest_gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 scripts/check_syncbn.py 2>&1 | tail -2 | tee gpurun_out/check_syncbn.log
for extra in "" "--sync-bn" "--graph" "--graph --sync-bn"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --config c3 --steps 10 --warmup 3 $extra 2>&1 | tail -1 | python -c "import json,sys; l=sys.stdin.read().strip().splitlines()[-1]
try:
    d=json.loads(l); print('c3 n=2 [$extra]', d['ms_per_step'], d['value'], d['config']['launch_mode'][:60])
except Exception as e: print('c3 n=2 [$extra] FAILED', l[:300])"
done
