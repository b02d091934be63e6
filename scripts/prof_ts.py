"""Profiling driver: a few forward calls of one narrow layer on indoor-shaped grids (run under ncu)."""
import sys

import torch

sys.path.insert(0, "fvdb-core_b200")
import fvdb
from fvdb import _fvdb_cpp as cpp
from fvdb.utils.synthetic import indoor_room

cin, cout, ks = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16, 16, 5)))
coords = [indoor_room(target=200_000, seed=10 + i, device="cuda") for i in range(8)]
grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(coords))
plan = fvdb.ConvolutionPlan.from_grid_batch(kernel_size=ks, stride=1, source_grid=grid, target_grid=grid)
topo = plan._backend.topology
n = grid.total_voxels
x = torch.randn((n, cin), device="cuda").bfloat16()
w = (torch.randn((cout, cin, ks, ks, ks), device="cuda") / (cin * ks**3) ** 0.5).bfloat16()
for _ in range(4):
    y = cpp.gs_conv(x, w, topo)
torch.cuda.synchronize()
print("ok", n, topo.total_pairs)
