"""Under torchrun (N ranks, NCCL): the by-grid partition of ONE batch reproduces the one-rank result.

Every rank builds the whole C2 batch (same seeds) and runs it alone (the reference result), then runs only its LPT share and
all-reduces grad_weights over NCCL -- the path's only exchange.  y / grad_x must match the whole-batch rows bit for bit (the map
never crosses grids, GatherScatterDefault.cu:126,186-188); the all-reduced grad_w must match to bf16 round-off.
"""
import os
import sys

sys.path.insert(0, "fvdb-core_b200")
sys.path.insert(0, ".")
import torch
import torch.distributed as dist

import bench
import fvdb
from fvdb import _fvdb_cpp as cpp

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = bench.CONFIGS["c2"]
coords = bench.make_coords(cfg, 0, dev)
whole = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(coords))
plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, whole, whole)
n = whole.total_voxels
gen = torch.Generator().manual_seed(7)
x = torch.randn((n, 64), generator=gen).bfloat16().to(dev)
w = ((torch.rand((64, 64, 3, 3, 3), generator=gen) * 2 - 1) / (64 * 27) ** 0.5).bfloat16().to(dev)
dy = torch.randn((n, 64), generator=gen).bfloat16().to(dev)
y1 = cpp.gs_conv(x, w, plan._backend.topology)
gx1, gw1 = cpp.gs_conv_backward(dy, x, w, plan._backend.topology)
sizes = [int(c.shape[0]) for c in coords]
offsets = [0]
for s in sizes:
    offsets.append(offsets[-1] + s)
mine = bench.partition_grids_lpt(sizes, world)[rank]
part = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor([coords[g] for g in mine]))
rows = torch.cat([torch.arange(offsets[g], offsets[g + 1], device=dev) for g in mine])
topo = fvdb.ConvolutionPlan.from_grid_batch(3, 1, part, part)._backend.topology
y = cpp.gs_conv(x[rows], w, topo)
gx, gw = cpp.gs_conv_backward(dy[rows], x[rows], w, topo)
dist.all_reduce(gw)  # SUM over ranks, in bf16 as the training step does
ok_rows = bool(torch.equal(y, y1[rows])) and bool(torch.equal(gx, gx1[rows]))
err = float((gw.float() - gw1.float()).abs().max() / gw1.float().abs().max())
flags = torch.tensor([float(ok_rows), err], device=dev)
gathered = [torch.zeros_like(flags) for _ in range(world)]
dist.all_gather(gathered, flags)
if rank == 0:
    print({"world": world, "grids_per_rank": [len(bench.partition_grids_lpt(sizes, world)[r]) for r in range(world)],
           "y_and_grad_x_bit_equal_on_every_rank": all(bool(g[0] > 0) for g in gathered), "grad_w_max_rel_err_after_allreduce": max(float(g[1]) for g in gathered)})
    assert all(bool(g[0] > 0) for g in gathered) and max(float(g[1]) for g in gathered) < 2e-2
dist.destroy_process_group()
