"""Times the plan-build kernels (kernel map, CSR view, tile mask, reverse map, grid build) on a bench config."""
import sys, json, time
sys.path.insert(0, "fvdb-core_b200"); sys.path.insert(0, ".")
import torch, fvdb, bench
from fvdb import _fvdb_cpp as cpp
from fvdb._lib import lib, check, i3
import ctypes as C

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = torch.device("cuda")
coords = bench.make_coords(cfg, 0, dev)
jt = fvdb.JaggedTensor(coords)
ev = lambda: torch.cuda.Event(enable_timing=True)
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = ev(), ev(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
grid = fvdb.GridBatch.from_ijk(jt)
k = cfg["kernel"]; ks = [k] * 3; st = [1] * 3
n, k3 = grid.total_voxels, k ** 3
res = {"voxels": n, "leaves": grid.total_leaf_nodes, "grid_build_ms(from_ijk, incl. host sync)": timed(lambda: fvdb.GridBatch.from_ijk(jt), 5)}
pitch = cpp._map_pitch(n)
nbr = torch.empty((k3, pitch), dtype=torch.int32, device=dev)
counts = torch.empty(k3, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
fused_mask = torch.empty(((n + 127) // 128) * ((k3 + 63) // 64), dtype=torch.int64, device=dev)
build = lambda: check(lib.fvc_kmap_build(grid.data.struct, grid.data.struct, i3(ks), i3(st), 0, nbr.data_ptr(), pitch, counts.data_ptr(), fused_mask.data_ptr(), stream))
t = timed(build)
wbytes = 4 * n * k3 + 128 * grid.total_leaf_nodes
res["kmap_build_ms(map + tap counts + tile mask, one kernel)"] = t; res["kmap_build_GBps(write 4*N*K3 + read 128 B/leaf)"] = wbytes / t / 1e6
res["kmap_build_frac_of_hbm_peak"] = wbytes / t / 1e6 / bench.load_peaks()["hbm_gbs"]
build_nomask = lambda: check(lib.fvc_kmap_build(grid.data.struct, grid.data.struct, i3(ks), i3(st), 0, nbr.data_ptr(), pitch, counts.data_ptr(), None, stream))
res["kmap_build_ms(without the fused tile mask)"] = timed(build_nomask)
topo = cpp.gs_build_topology(grid.data, grid.data, ks, st)
P = topo.total_pairs
gather = torch.empty(P, dtype=torch.int32, device=dev); scatter = torch.empty(P, dtype=torch.int32, device=dev)
offs = torch.empty(k3 + 1, dtype=torch.int64, device=dev)
sb = int(lib.fvc_kmap_csr_scratch_bytes(n, k3)); scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
t = timed(lambda: check(lib.fvc_kmap_to_csr(nbr.data_ptr(), pitch, n, k3, counts.data_ptr(), offs.data_ptr(), gather.data_ptr(), scatter.data_ptr(), scratch.data_ptr(), sb, stream)))
res["csr_ms"] = t; res["csr_GBps(2 reads of map + 8P written)"] = (8 * n * k3 + 8 * P) / t / 1e6
mask = torch.empty(((n + 127) // 128) * ((k3 + 63) // 64), dtype=torch.int64, device=dev)
t = timed(lambda: check(lib.fvc_kmap_tile_mask(nbr.data_ptr(), pitch, n, k3, mask.data_ptr(), stream)))
res["tile_mask_ms"] = t; res["tile_mask_GBps"] = 4 * n * k3 / t / 1e6
rev = torch.empty((k3, pitch), dtype=torch.int32, device=dev)
t = timed(lambda: check(lib.fvc_kmap_reverse_dense(gather.data_ptr(), scatter.data_ptr(), offs.data_ptr(), k3, P, n, rev.data_ptr(), pitch, stream)))
res["reverse_dense_ms(from CSR)"] = t
t = timed(lambda: check(lib.fvc_kmap_reverse_from_dense(nbr.data_ptr(), pitch, n, k3, n, rev.data_ptr(), pitch, stream)))
res["reverse_dense_ms(from the dense map)"] = t
assert torch.equal(mask, fused_mask), "fused tile mask differs from the stand-alone pass"
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    plan = fvdb.ConvolutionPlan.from_grid_batch(k, 1, grid, grid); plan._backend.topology._dgrad_plan()
torch.cuda.synchronize(); res["plan_total_ms(python, no host sync inside)"] = (time.perf_counter() - t0) / 5 * 1e3
torch.cuda.empty_cache(); torch.cuda.synchronize(); t0 = time.perf_counter()
plan = fvdb.ConvolutionPlan.from_grid_batch(k, 1, grid, grid); plan._backend.topology._dgrad_plan()
torch.cuda.synchronize(); res["plan_cold_ms(allocator cache emptied first)"] = (time.perf_counter() - t0) * 1e3
res["pairs"] = P
print(json.dumps(res))
