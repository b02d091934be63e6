"""Times the forward tcgen05 kernel for each pipeline-shape variant (fvc_set_tuning key 0) on a bench workload.

    python scripts/bench_variants.py [config=c2] [variants=0,1,2,...] [out.json]

variant 1 = the round-1 kernel (four producer warps sharing every unit + map ring); the others are warp-per-unit
shapes (csrc/conv_tc.cu: tc_forward_half).  Every variant's output is compared with variant 1's (bit-equal expected:
same MMA order per accumulator).  L2 is flushed between timed launches? No -- inputs (features + map) exceed L2.
"""
import json
import sys

sys.path.insert(0, "fvdb-core_b200")
sys.path.insert(0, ".")
import torch

import bench
import fvdb
from fvdb import _fvdb_cpp as cpp

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
variants = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "1,0,2,3,4,5,6").split(",")]
dev = torch.device("cuda")
coords = bench.make_coords(cfg, 0, dev)
grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(coords))
plan = fvdb.ConvolutionPlan.from_grid_batch(cfg["kernel"], 1, grid, grid)
topo = plan._backend.topology
n, k3, cin, cout = grid.total_voxels, topo.kernel_volume, cfg["cin"], cfg["cout"]
dtype = bench.DTYPES[cfg["dtype"]]
x = torch.randn((n, cin), device=dev).to(dtype)
w = (torch.randn((cout, cin, cfg["kernel"], cfg["kernel"], cfg["kernel"]), device=dev) * 0.02).to(dtype)
wp = cpp._prepare_weights(w, dtype, False)
ref, rows = None, []
for variant in variants:
    cpp.set_kernel_variant(variant)
    f = lambda: cpp._run_conv(x, wp, topo._out_map(), n, n, cin, cout, k3, None, topo._out_mask())  # noqa: E731
    try:
        y = f()
        torch.cuda.synchronize()
    except RuntimeError as exc:
        rows.append({"variant": variant, "error": str(exc)[:200]})
        print(json.dumps(rows[-1]), flush=True)
        continue
    if ref is None:
        ref = y
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        f()
    b.record()
    torch.cuda.synchronize()
    rows.append({"config": cfg["desc"], "variant": variant, "fwd_ms": a.elapsed_time(b) / 10, "max_abs_diff_vs_first": float((y.float() - ref.float()).abs().max())})
    print(json.dumps(rows[-1]), flush=True)
cpp.set_kernel_variant(0)
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as fh:
        json.dump(rows, fh, indent=1)
# weight-gradient shapes (key 1) on the same batch
if cin == cout and cin in (64, 128, 256):
    dy = torch.randn((n, cout), device=dev).to(dtype)
    wrows, wref = [], None
    # 3 = two pipelines (64) / the single-pipeline baseline (128, 256); 5 = contiguous tile ranges per CTA, 6 / 7 / 8 = mini-chunks of 8 / 32 / 64 tiles
    for variant in ((0, 1, 2, 3, 5, 6, 7, 8) if cin == 64 else (0, 3, 5, 6, 7, 8)):
        cpp.set_kernel_variant(variant, wgrad=True)
        f = lambda: cpp.gs_conv_backward(dy, x, w, topo, need_grad_features=False)[1]  # noqa: E731
        gw = f()
        torch.cuda.synchronize()
        wref = gw if wref is None else wref
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            f()
        b.record()
        torch.cuda.synchronize()
        wrows.append({"wgrad_variant": variant, "wgrad_ms": a.elapsed_time(b) / 10, "max_abs_diff_vs_first": float((gw.float() - wref.float()).abs().max()),
                      "max_abs": float(wref.float().abs().max())})
        print(json.dumps(wrows[-1]), flush=True)
    cpp.set_kernel_variant(0, wgrad=True)
    if len(sys.argv) > 3:
        with open(sys.argv[3].replace(".json", "_wgrad.json"), "w") as fh:
            json.dump(wrows, fh, indent=1)
