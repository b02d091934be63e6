#!/bin/bash
# 2-GPU call: default bench under torchrun (N=2), C4 on one GPU (32 grids x 1M), ncu full capture of the C5 forward kernel.
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -2 | tee gpurun_out/bench_n2.json | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_n2_reference.json | cut -c1-600
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_c4.json | cut -c1-3000
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_fwd|conv_tc_wgrad" -s 2 -c 2 -o gpurun_out/prof_c5 -f python bench.py --config c5 --grids 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c5_under_ncu.log 2>&1
tail -1 gpurun_out/c5_under_ncu.log | cut -c1-300
