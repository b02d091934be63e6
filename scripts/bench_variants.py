"""Times the forward tcgen05 kernel on the C2 workload for each FVC_TC_VARIANT (pipeline-shape experiment)."""
import os, sys, json
sys.path.insert(0, "fvdb-core_b200"); sys.path.insert(0, ".")
import torch, fvdb, bench
from fvdb import _fvdb_cpp as cpp

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = torch.device("cuda")
coords = bench.make_coords(cfg, 0, dev)
grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(coords))
plan = fvdb.ConvolutionPlan.from_grid_batch(cfg["kernel"], 1, grid, grid)
topo = plan._backend.topology
n, k3, cin, cout = grid.total_voxels, topo.kernel_volume, cfg["cin"], cfg["cout"]
dtype = bench.DTYPES[cfg["dtype"]]
x = torch.randn((n, cin), device=dev).to(dtype)
w = (torch.randn((cout, cin, cfg["kernel"], cfg["kernel"], cfg["kernel"]), device=dev) * 0.02).to(dtype)
wp = cpp._pack_weights(w, dtype, 0)
ref = None
for variant in [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else "0,1,2,3,10".split(","))]:
    os.environ["FVC_TC_VARIANT"] = str(variant)
    f = lambda: cpp._run_conv(x, wp, topo._out_map(), n, n, cin, cout, k3, None, topo._out_mask())
    y = f(); torch.cuda.synchronize()
    if ref is None: ref = y
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): f()
    b.record(); torch.cuda.synchronize()
    print(json.dumps({"variant": variant, "fwd_ms": a.elapsed_time(b) / 5, "max_abs_diff_vs_v0": float((y.float() - ref.float()).abs().max())}), flush=True)
