#!/bin/bash
# N-GPU call: default bench (weak scaling, C2 per rank); optionally C4 partitioned by grid (strong scaling) with "c4" as 2nd arg.
set -u
mkdir -p gpurun_out
N=${1:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n$N.json | cut -c1-700
if [ "${2:-}" = "c4" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --config c4 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c4_n$N.json | cut -c1-700
fi
