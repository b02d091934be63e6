#!/bin/bash
# 2-GPU check: default bench under torchrun (weak scaling, asynchronous all-reduce behind wgrad), reference arm, SyncBatchNorm check.
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n$N.json | cut -c1-900
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 scripts/check_syncbn.py 2>&1 | tail -2 | cut -c1-200
