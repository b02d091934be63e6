"""Measurement of the SURVEY.md 8(f) rows around the convolution (grid build, generated grids, lookups, normalisation,
pooling) on the default bench batch: CUDA-event times and achieved GB/s over each kernel group's algorithmic bytes,
against the measured HBM peak.  Prints one JSON object.  python scripts/bench_next.py [config]"""
import json, sys
sys.path.insert(0, "fvdb-core_b200"); sys.path.insert(0, ".")
import torch, fvdb, bench

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = torch.device("cuda")
peak = bench.load_peaks()["hbm_gbs"]
coords = bench.make_coords(cfg, 0, dev)
jt = fvdb.JaggedTensor(coords)
ev = lambda: torch.cuda.Event(enable_timing=True)

def timed(fn, reps=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = ev(), ev(); a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def row(ms, nbytes):
    return {"ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / ms / 1e6, 1), "frac_of_hbm_peak": round(nbytes / ms / 1e6 / peak, 3)}

grid = fvdb.GridBatch.from_ijk(jt)
n, leaves = grid.total_voxels, grid.total_leaf_nodes
n_in = sum(int(c.shape[0]) for c in coords)
res = {"workload": cfg["desc"], "voxels": n, "leaves": leaves, "hbm_peak_GBps": peak}
# grid build: 12 B in + 16 B out per coordinate + 56 B of sort scratch traffic; includes its one D2H count sync
res["from_ijk (sort + node fill, incl. host sync)"] = row(timed(lambda: fvdb.GridBatch.from_ijk(jt), 5), n_in * (12 + 16 + 56))
pts = fvdb.JaggedTensor([c.float() * 0.02 + 0.003 for c in coords])
res["from_points (voxelise + from_ijk)"] = row(timed(lambda: fvdb.GridBatch.from_points(pts, 0.02), 5), n_in * (12 + 12 + 12 + 16 + 56))
down = grid.conv_grid(2, 2)
res["conv_grid k=2 s=2 (candidates + build)"] = row(timed(lambda: grid.conv_grid(2, 2), 5), n * 16 + n * (12 + 16 + 56))
dil = grid.conv_grid(3, 1)
res["conv_grid k=3 s=1 (27 candidates / voxel + build)"] = row(timed(lambda: grid.conv_grid(3, 1), 5), 27 * n * (16 + 12 + 56) + dil.total_voxels * 16)
res["conv_grid outputs"] = {"k2s2_voxels": down.total_voxels, "k3s1_voxels": dil.total_voxels}
k = 3
res["plan 3^3 (map + CSR + tile masks + reversed map, python incl. one sync)"] = row(
    timed(lambda: fvdb.ConvolutionPlan.from_grid_batch(k, 1, grid, grid)._backend.topology._dgrad_plan(), 5), 128 * leaves + 4 * n * 27 * 2 * 3)
q = grid.ijk
res["neighbor_indexes extent=1 (27 lookups / voxel, int64 out)"] = row(timed(lambda: grid.neighbor_indexes(q, 1)), n * (12 + 27 * 8))
res["ijk_to_index (1 lookup / voxel)"] = row(timed(lambda: grid.ijk_to_index(q)), n * (12 + 8))
for c, dt in ((64, torch.bfloat16), (32, torch.float32)):
    x = torch.randn((n, c), device=dev).to(dt).requires_grad_()
    dy = torch.randn((n, c), device=dev).to(dt)
    s = x.element_size()
    bn = fvdb.nn.BatchNorm(c, activation="relu").to(dev)
    tbn = torch.nn.BatchNorm1d(c).to(dev)
    name = f"[{c} ch {str(dt).split('.')[-1]}]"
    res[f"BatchNorm+ReLU forward {name} (3 passes)"] = row(timed(lambda: bn(grid.jagged_like(x), grid)), 3 * n * c * s)
    y = bn(grid.jagged_like(x), grid).jdata
    res[f"BatchNorm+ReLU backward {name} (5 passes)"] = row(timed(lambda: torch.autograd.grad(y, x, dy, retain_graph=True)), 5 * n * c * s)
    res[f"torch BatchNorm1d + relu forward {name} (same bytes)"] = row(timed(lambda: torch.relu(tbn(x))), 3 * n * c * s)
    yt = torch.relu(tbn(x))
    res[f"torch BatchNorm1d + relu backward {name} (same bytes)"] = row(timed(lambda: torch.autograd.grad(yt, x, dy, retain_graph=True)), 5 * n * c * s)
    coarse = grid.coarsened_grid(2)
    res[f"max_pool 2 forward {name} (child table + rows)"] = row(timed(lambda: grid.max_pool(2, grid.jagged_like(x), coarse_grid=coarse)),
                                                                 n * c * s + coarse.total_voxels * (c * s + 32 + 12 + 8 * 8))
    p, _ = grid.max_pool(2, grid.jagged_like(x), coarse_grid=coarse)
    dp = torch.randn_like(p.jdata)
    res[f"max_pool 2 backward {name}"] = row(timed(lambda: torch.autograd.grad(p.jdata, x, dp, retain_graph=True)), 2 * n * c * s + coarse.total_voxels * (c * s + 32))
    z = torch.randn((coarse.total_voxels, c), device=dev).to(dt)
    res[f"refine 2 forward {name} (parent table + rows)"] = row(timed(lambda: coarse.refine(2, coarse.jagged_like(z), fine_grid=grid)), n * (c * s + 4 + 12 + 8) + coarse.total_voxels * c * s)
print(json.dumps(res))
