#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_c3.json | cut -c1-1300
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --graph 2>&1 | tail -3 | tee gpurun_out/bench_c3_graph.json | cut -c1-1300
timeout 600 python bench.py --config c3 --steps 10 --warmup 3 --torch-bn 2>&1 | tail -1 | tee gpurun_out/bench_c3_torchbn.json | cut -c1-300
