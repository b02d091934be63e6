"""Small forward + backward calls of every tensor-core kernel organisation of round 2 (run under compute-sanitizer memcheck):
pipelined CTAs (128 / 256 wide), fused backward and tensor-memory executor (narrow), 1-tile fp32 CTAs, mini-chunk weight gradient."""
import sys

import numpy as np
import torch

sys.path.insert(0, "fvdb-core_b200")
import fvdb
from fvdb import _fvdb_cpp as cpp
from fvdb.utils.synthetic import sphere_shell

shell = sphere_shell(target=5000, domain=64, seed=3, device="cpu").numpy()
grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor([torch.tensor(shell.astype(np.int32), device="cuda")]))
n = grid.total_voxels
for dtype, cin, cout, ks, variant in [(torch.bfloat16, 128, 128, 3, 0), (torch.bfloat16, 256, 256, 3, 0), (torch.bfloat16, 64, 64, 3, 0), (torch.bfloat16, 32, 32, 3, 0),
                                      (torch.bfloat16, 16, 16, 5, 0), (torch.bfloat16, 16, 32, 3, 12), (torch.float32, 32, 32, 3, 0), (torch.float32, 64, 64, 3, 0),
                                      (torch.bfloat16, 64, 64, 3, 13)]:
    plan = fvdb.ConvolutionPlan.from_grid_batch(kernel_size=ks, stride=1, source_grid=grid, target_grid=grid)
    topo = plan._backend.topology
    x = torch.randn((n, cin), device="cuda").to(dtype)
    w = (torch.randn((cout, cin, ks, ks, ks), device="cuda") / (cin * ks**3) ** 0.5).to(dtype)
    dy = torch.randn((n, cout), device="cuda").to(dtype)
    cpp.set_kernel_variant(variant)
    y = cpp.gs_conv(x, w, topo)
    gx, gw = cpp.gs_conv_backward(dy, x, w, topo)
    cpp.set_kernel_variant(0)
    torch.cuda.synchronize()
    print(str(dtype).split(".")[-1], cin, cout, ks, "variant", variant, "ok", float(y.float().abs().mean()), float(gw.float().abs().mean()), flush=True)
