#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 10 11 12; do
echo "== variant $v"
FVC_TC_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "identity_map_is_plain_gemm and 64-64 or (values_and_gradients and bfloat16-64-64-3-1) or determinism" 2>&1 | tail -4
FVC_TC_VARIANT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 variant $v', d['ms_per_step'], {k:(round(v['ms'],4), round(v['frac'],3)) for k,v in d['roofline_kernels'].items()})"
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c2 default', d['ms_per_step'], {k:(round(v['ms'],4), round(v['frac'],3)) for k,v in d['roofline_kernels'].items()})"
