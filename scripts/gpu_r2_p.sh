#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/bench_variants.py c2 0 gpurun_out/r2p_variants_c2.json 2>&1 | grep -E "^\{|rror" | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_memory_executor" 2>&1 | tail -3
