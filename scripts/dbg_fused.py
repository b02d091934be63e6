"""Debug driver: one fused-backward call on a small batch (run with FVC_DEBUG_WAITS=1, optionally under compute-sanitizer)."""
import sys

import numpy as np
import torch

sys.path.insert(0, "fvdb-core_b200")
import fvdb
from fvdb import _fvdb_cpp as cpp
from fvdb.utils.synthetic import sphere_shell

cin, cout, ks = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (16, 16, 5)))
target = int(sys.argv[4]) if len(sys.argv) > 4 else 3000
shell = sphere_shell(target=target, domain=64, seed=3, device="cpu").numpy()
grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor([torch.tensor(shell.astype(np.int32), device="cuda")]))
plan = fvdb.ConvolutionPlan.from_grid_batch(kernel_size=ks, stride=1, source_grid=grid, target_grid=grid)
topo = plan._backend.topology
n = grid.total_voxels
gen = torch.Generator().manual_seed(1)
x = torch.randn((n, cin), generator=gen).bfloat16().cuda()
w = (torch.randn((cout, cin, ks, ks, ks), generator=gen) / (cin * ks ** 3) ** 0.5).bfloat16().cuda()
dy = torch.randn((n, cout), generator=gen).bfloat16().cuda()
torch.cuda.synchronize()
print("rows", n, "family", cpp.lib.fvc_conv_kernel_family(cin, cout, ks ** 3, cpp._DTYPE_CODE[torch.bfloat16], 0, 2), flush=True)
gx, gw = cpp.gs_conv_backward(dy, x, w, topo)
torch.cuda.synchronize()
cpp.set_fused_backward(False)
gx2, gw2 = cpp.gs_conv_backward(dy, x, w, topo)
torch.cuda.synchronize()
print("dx rel err", float((gx.float() - gx2.float()).norm() / gx2.float().norm()), "dw rel err", float((gw.float() - gw2.float()).norm() / gw2.float().norm()))
