"""Experiment: how precise is the tcgen05 fp32 accumulation over long chains?  Reads the fp32 partial slices of
the bf16 tensor-core wgrad kernel (products of bf16 values are exact in fp32, so any deviation from an fp64
reference is accumulation error) for random-sign and all-positive inputs."""
import sys, json
sys.path.insert(0, "fvdb-core_b200"); sys.path.insert(0, ".")
import ctypes as C
import torch, fvdb, bench
from fvdb import _fvdb_cpp as cpp
from fvdb._lib import lib, check

dev = torch.device("cuda")
cfg = dict(bench.CONFIGS["c2"]); cfg["grids"] = 2
coords = bench.make_coords(cfg, 0, dev)
grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(coords))
plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid, grid)
topo = plan._backend.topology
n, k3, cin, cout = grid.total_voxels, 27, 64, 64
res = {"voxels": n, "pairs": topo.total_pairs}
for mode in ("randn", "positive"):
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randn((n, cin), generator=g, device=dev)
    dy = torch.randn((n, cout), generator=g, device=dev)
    if mode == "positive":
        x, dy = x.abs(), dy.abs()
    x, dy = x.bfloat16(), dy.bfloat16()
    code = cpp._DTYPE_CODE[torch.bfloat16]
    sb = int(lib.fvc_conv_wgrad_scratch_bytes(n, n, topo.total_pairs, cin, cout, k3, code, 0, 1))
    scratch = torch.zeros(sb, dtype=torch.uint8, device=dev)
    gw = torch.empty((cout, cin, 3, 3, 3), dtype=torch.bfloat16, device=dev)
    out_map = topo._out_map()
    check(lib.fvc_conv_wgrad(x.data_ptr(), dy.data_ptr(), topo.gather_indices.data_ptr(), topo.scatter_indices.data_ptr(),
                             C.cast(topo.offsets.data_ptr(), C.POINTER(C.c_int64)), topo._core.offsets_dev.data_ptr(), out_map.data_ptr(), int(out_map.shape[1]),
                             topo._out_mask().data_ptr(), n, n, cin, cout, k3, code, 2, gw.data_ptr(), scratch.data_ptr(), sb, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    slice_elems = k3 * cin * cout
    chunks = (sb - 256) // (slice_elems * 4)
    partial = scratch[: chunks * slice_elems * 4].view(torch.float32).view(chunks, k3, cin, cout)
    got = partial.double().sum(0)
    ref = torch.zeros((k3, cin, cout), dtype=torch.float64, device=dev)
    offs = topo.offsets.tolist()
    xd, dd = x.double(), dy.double()
    for k in range(k3):
        gi, si = topo.gather_indices[offs[k]:offs[k + 1]].long(), topo.scatter_indices[offs[k]:offs[k + 1]].long()
        ref[k] = xd[gi].T @ dd[si]
    rel = float((got - ref).norm() / ref.norm())
    ratio = ((got - ref) / ref.abs().clamp_min(1e-30))
    res[mode] = {"chunks": int(chunks), "rel_err_norm": rel, "mean_signed_rel": float(ratio[ref.abs() > ref.abs().mean() * 0.1].mean()),
                 "tiles_per_chunk": (n + 127) // 128 / (chunks), "mma_steps_per_accumulator~": (n + 127) // 128 / chunks * 8 * 0.5}
print(json.dumps(res))
