"""Where does ConvolutionPlan.from_grid_batch spend host time?  cProfile over 20 builds on the C2 batch."""
import sys, cProfile, pstats, io, time
sys.path.insert(0, "fvdb-core_b200"); sys.path.insert(0, ".")
import torch, fvdb, bench
cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = torch.device("cuda")
coords = bench.make_coords(cfg, 0, dev)
jt = fvdb.JaggedTensor(coords)
grid = fvdb.GridBatch.from_ijk(jt)
k = cfg["kernel"]
def build():
    plan = fvdb.ConvolutionPlan.from_grid_batch(k, 1, grid, grid)
    plan._backend.topology._dgrad_plan()
    return plan
for _ in range(3): build()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): build()
torch.cuda.synchronize()
print("plan build ms (mean of 20):", (time.perf_counter() - t0) / 20 * 1e3)
t0 = time.perf_counter()
for _ in range(10): fvdb.GridBatch.from_ijk(jt)
torch.cuda.synchronize()
print("from_ijk ms (mean of 10):", (time.perf_counter() - t0) / 10 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(20): build()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
