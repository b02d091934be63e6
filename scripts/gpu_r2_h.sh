#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/time_fused.py 2>&1 | grep -E "^\{|Error|error" | tee gpurun_out/r2h_time_fused.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_backward" 2>&1 | tail -4
