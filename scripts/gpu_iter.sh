#!/bin/bash
set -u
mkdir -p gpurun_out
for lib in fvdb-core_b200/fvdb/libfvdbconv.so fvdb-core_b200/fvdb/libfvdbconv_bn1.so fvdb-core_b200/fvdb/libfvdbconv.so fvdb-core_b200/fvdb/libfvdbconv_bn1.so; do
FVC_LIB=$PWD/$lib timeout 300 python scripts/bench_next.py c2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', {k.split('(')[0][15:]:v['ms'] for k,v in d.items() if k.startswith('BatchNorm+ReLU')})"
done
timeout 300 python scripts/bench_variants.py c2x128 0,5,6 2>&1 | tail -3
timeout 300 python scripts/bench_variants.py c1 0,5,6 2>&1 | tail -3
timeout 300 python scripts/bench_variants.py c3 0,5,6 2>&1 | tail -3
