"""Fused backward (one gather) against the separate dgrad + wgrad kernels on indoor-shaped grids: ms per backward pass."""
import json
import sys

import torch

sys.path.insert(0, "fvdb-core_b200")
sys.path.insert(0, ".")
import fvdb
from fvdb import _fvdb_cpp as cpp
from fvdb.utils.synthetic import indoor_room


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = []
coords = [indoor_room(target=200_000, seed=10 + i, device="cuda") for i in range(8)]
grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(coords))
for cin, cout, ks in [(32, 32, 3), (16, 16, 3), (16, 16, 5), (32, 16, 3), (16, 32, 3)]:
    plan = fvdb.ConvolutionPlan.from_grid_batch(kernel_size=ks, stride=1, source_grid=grid, target_grid=grid)
    topo = plan._backend.topology
    n = grid.total_voxels
    x = torch.randn((n, cin), device="cuda").bfloat16()
    dy = torch.randn((n, cout), device="cuda").bfloat16()
    w = (torch.randn((cout, cin, ks, ks, ks), device="cuda") / (cin * ks**3) ** 0.5).bfloat16()
    rec = {"shape": f"{cin}->{cout} {ks}^3", "voxels": n}
    for name, flag in (("fused_ms", True), ("separate_ms", False)):
        cpp.set_fused_backward(flag)
        rec[name] = round(timed(lambda: cpp.gs_conv_backward(dy, x, w, topo)), 4)
    rec["fwd_ms"] = round(timed(lambda: cpp.gs_conv(x, w, topo)), 4)
    out.append(rec)
    print(json.dumps(rec), flush=True)
