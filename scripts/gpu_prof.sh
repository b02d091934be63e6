#!/bin/bash
# ncu captures: launch list of a short bench run + full sections of the three hot kernels.
set -u
mkdir -p gpurun_out
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/mask_stats.log
import sys; sys.path.insert(0,'fvdb-core_b200'); sys.path.insert(0,'.')
import torch, fvdb, bench
cfg = bench.CONFIGS['c2']
coords = bench.make_coords(cfg, 0, torch.device('cuda'))
grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(coords))
plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid, grid)
topo = plan._backend.topology
m = topo._out_mask()
bits = sum(int(((m >> k) & 1).sum()) for k in range(27))
print('tiles', m.numel(), 'active units', bits, 'fraction', bits / (m.numel()*27), 'pairs', topo.total_pairs, 'fill of active units', topo.total_pairs/(bits*128))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_fwd|conv_tc_wgrad" -s 2 -c 3 -o gpurun_out/prof_tc -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out/
