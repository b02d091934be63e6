#!/bin/bash
# e2e (host-buffer) pipeline check: parity test of HostPipelinedConv, then the bench line for two chunk counts
set -u
timeout 200 python -m pytest tests -m gpu -x -q -k "host_pipelined" 2>&1 | grep -E "^E  |passed|failed|Error" | head -8
for c in 8 16; do
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-chunks $c 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunks $c', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['host_link_probe'])"
done
