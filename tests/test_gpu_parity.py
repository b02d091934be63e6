"""GPU parity tests: the CUDA path (through the C ABI, via the fvdb mirror) against the CPU oracle and the
golden vectors produced by the reference's own oracle.  Bit-exact for every integer product (grids, kernel
maps, neighbour indices); fp64 1e-11 (the reference's own bar, test_conv_semantics_integration.py:214-242);
fp32 1e-5 relative; f16/bf16 2e-2 against the fp32 oracle (north_star).
"""

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fvdb():
    import fvdb as module

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return module


DEV = "cuda"


def _grid(fvdb, coords_per_grid, **kw):
    tensors = [torch.tensor(np.asarray(c, dtype=np.int32).reshape(-1, 3), device=DEV) for c in coords_per_grid]
    return fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(tensors), **kw)


def _rows(grid):
    return grid.ijk.jdata.cpu().numpy().astype(np.int64), grid.jidx.cpu().numpy().astype(np.int64)


def _random_batch(seed, n=3000, extent=40, batches=3, dup=True):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(batches):
        c = rng.integers(-extent, extent, size=(n, 3))
        if dup:
            c = np.concatenate([c, c[: n // 10]])
        out.append(c)
    return out


def _canonical(gather, scatter, offsets):
    """Sort every tap segment by (output row, feature row): the reference leaves pair order unspecified."""
    gather, scatter = np.asarray(gather).astype(np.int64), np.asarray(scatter).astype(np.int64)
    g2, s2 = gather.copy(), scatter.copy()
    for k in range(len(offsets) - 1):
        a, b = int(offsets[k]), int(offsets[k + 1])
        order = np.lexsort((gather[a:b], scatter[a:b]))
        g2[a:b], s2[a:b] = gather[a:b][order], scatter[a:b][order]
    return g2, s2


# ------------------------------------------------------------------ index grid


def test_grid_build_rows_order_and_lookup(fvdb):
    coords = _random_batch(1)
    coords[1] = np.zeros((0, 3), dtype=np.int64)  # empty batch item is preserved
    coords.append(np.array([[5000, -9000, 4096], [-4097, 0, 1], [2**20, -(2**20), 77], [5000, -9000, 4096]]))  # several root tiles
    grid = _grid(fvdb, coords)
    ijk, b = _rows(grid)
    want = np.unique(np.concatenate([np.concatenate([np.full((len(c), 1), i), c], axis=1) for i, c in enumerate(coords)]), axis=0)
    order = oracle.index_grid_row_order(want[:, 0], want[:, 1:])
    assert grid.total_voxels == len(want) and grid.grid_count == 4
    assert np.array_equal(b, want[order, 0]) and np.array_equal(ijk, want[order, 1:])  # bit-exact incl. row order
    assert grid.joffsets.cpu().tolist() == np.concatenate([[0], np.cumsum(np.bincount(want[:, 0], minlength=4))]).tolist()
    # ijk_to_index o ijk == identity (reference tests/unit/test_basic_ops.py), misses are -1
    idx = grid.ijk_to_index(grid.ijk, cumulative=True).jdata.cpu().numpy()
    assert np.array_equal(idx, np.arange(len(want)))
    miss = fvdb.JaggedTensor([torch.tensor([[999, 999, 999]], dtype=torch.int32, device=DEV)] * 4)
    assert grid.ijk_to_index(miss).jdata.cpu().tolist() == [-1] * 4
    assert not bool(grid.coords_in_grid(miss).jdata.any())


def test_from_points_voxelises_like_the_reference_transform(fvdb):
    # ijk = round(p / voxel_size - origin / voxel_size), round = floor(x + 0.5)  (VoxelCoordTransform.h:300-310,
    # BuildGridFromPoints.cu:89); per-grid voxel sizes and origins, negative coordinates, duplicate voxels merged.
    gen = torch.Generator().manual_seed(17)
    pts = [torch.randn((4000, 3), generator=gen) * 3.0, torch.rand((2500, 3), generator=gen) * 5.0 - 4.0]
    sizes = torch.tensor([[0.25, 0.5, 0.125], [0.1, 0.1, 0.2]], dtype=torch.float64)
    origins = torch.tensor([[0.05, -0.3, 0.0], [1.0, 2.0, -3.0]], dtype=torch.float64)
    grid = fvdb.GridBatch.from_points(fvdb.JaggedTensor([p.to(DEV) for p in pts]), sizes, origins)
    for b, p in enumerate(pts):
        want = np.unique(np.floor(p.numpy().astype(np.float32) * (1.0 / sizes[b].numpy()).astype(np.float32)
                                  + (-origins[b].numpy() / sizes[b].numpy()).astype(np.float32) + np.float32(0.5)).astype(np.int64), axis=0)
        lo, hi = int(grid.joffsets[b]), int(grid.joffsets[b + 1])
        got = grid.ijk.jdata[lo:hi].cpu().numpy().astype(np.int64)
        assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, want.tolist()))
    torch.testing.assert_close(grid.voxel_sizes.cpu().double(), sizes)
    torch.testing.assert_close(grid.origins.cpu().double(), origins)
    with pytest.raises(TypeError):
        fvdb.GridBatch.from_points(torch.zeros((4, 3), dtype=torch.int32, device=DEV))


def test_neighbor_indexes_match_oracle(fvdb):
    grid = _grid(fvdb, _random_batch(2, n=2000, extent=12))
    ijk, b = _rows(grid)
    offsets = grid.joffsets.cpu().numpy()
    for extent, shift in ((1, 0), (2, 0), (1, 1)):
        got = grid.neighbor_indexes(grid.ijk, extent, shift).jdata.cpu().numpy()
        assert np.array_equal(got, oracle.neighbor_indexes(ijk, b, offsets, ijk, b, extent, shift))


# ------------------------------------------------------------------ generated topologies + kernel maps

_GEOMETRIES = [((k, k, k), (s, s, s)) for k in range(1, 7) for s in range(1, 6)] + [
    ((2, 3, 4), (1, 2, 3)), ((5, 2, 3), (4, 2, 1)), ((3, 4, 2), (2, 3, 4)), ((3, 5, 7), (1, 1, 1)), ((7, 7, 7), (1, 1, 1)), ((1, 1, 1), (3, 2, 1))]


@pytest.mark.parametrize("ks,st", _GEOMETRIES)
def test_generated_grids_and_kernel_maps(fvdb, ks, st):
    if ks == (1, 1, 1) and st == (1, 1, 1):
        pytest.skip("identity geometry returns the same grid object")
    coords = [[(-4, -1, 0), (-1, 0, 1), (0, 2, -3), (3, -2, 4), (5, 1, -1)], [], _random_batch(3, n=300, extent=9, batches=1)[0]]
    fine = _grid(fvdb, coords)
    f_ijk, f_b = _rows(fine)
    coarse = fine.conv_grid(ks, st)
    c_ijk, c_b = _rows(coarse)
    want_ijk, want_b = oracle.conv_grid(f_ijk, f_b, ks, st)
    order = oracle.index_grid_row_order(want_b, want_ijk)
    assert np.array_equal(c_ijk, want_ijk[order]) and np.array_equal(c_b, want_b[order])
    back = coarse.conv_transpose_grid(ks, st)
    bk_ijk, bk_b = _rows(back)
    want_ijk, want_b = oracle.conv_transpose_grid(c_ijk, c_b, ks, st)
    order = oracle.index_grid_row_order(want_b, want_ijk)
    assert np.array_equal(bk_ijk, want_ijk[order]) and np.array_equal(bk_b, want_b[order])
    torch.testing.assert_close(coarse.voxel_sizes.cpu(), torch.tensor([st] * 3, dtype=torch.float32))
    torch.testing.assert_close(back.voxel_sizes.cpu(), fine.voxel_sizes.cpu())

    cpp = fvdb._fvdb_cpp
    for transposed, (feat, out) in ((False, (fine, coarse)), (True, (coarse, fine))):
        build = cpp.gs_build_transpose_topology if transposed else cpp.gs_build_topology
        topo = build(feat.data, out.data, list(ks), list(st))
        ref = oracle.build_topology(*_rows(feat), *_rows(out), ks, st, transposed=transposed)
        assert topo.offsets.tolist() == ref.offsets.tolist() and topo.total_pairs == ref.total_pairs
        assert topo.gather_indices.dtype == torch.int32 and topo.offsets.device.type == "cpu" and topo.is_transposed == transposed
        g, s = _canonical(topo.gather_indices.cpu().numpy(), topo.scatter_indices.cpu().numpy(), ref.offsets)
        rg, rs = _canonical(ref.gather_indices, ref.scatter_indices, ref.offsets)
        assert np.array_equal(g, rg) and np.array_equal(s, rs)  # bit-exact kernel map
        dense = topo._out_map().cpu().numpy()[:, : out.total_voxels].T
        assert np.array_equal(dense, oracle.dense_kernel_map(*_rows(feat), *_rows(out), ks, st, transposed))
        rev = cpp.gs_reverse_topology(topo)
        assert rev.gather_indices.data_ptr() == topo.scatter_indices.data_ptr() and rev.is_transposed != transposed
        want_rev = oracle.dense_kernel_map(*_rows(out), *_rows(feat), ks, st, not transposed)
        assert np.array_equal(rev._out_map().cpu().numpy()[:, : feat.total_voxels].T, want_rev)


def test_topology_validator_accepts_built_maps_and_rejects_corrupted_ones(fvdb):
    # validateGatherScatterDefaultTopology (GatherScatterDefault.cu:342-527) as exercised by
    # src/tests/GatherScatterDefaultConvTest.cu:832-930: forward, reversed and round-trip views validate; corrupted metadata,
    # offsets, index ranges and non-canonical edges are rejected
    cpp = fvdb._fvdb_cpp
    fine = _grid(fvdb, [[(-5, -2, 0), (-1, 0, 2), (0, 1, 2), (3, 4, 5), (7, 2, -3)]])
    ks, st = (4, 3, 2), (3, 2, 4)
    coarse = cpp.conv_grid(fine.data, ks, st)
    forward = cpp.gs_build_topology(fine.data, coarse, ks, st)
    cpp.validate_gather_scatter_default_topology(fine.data, coarse, forward)
    reverse = cpp.gs_reverse_topology(forward)
    assert reverse.is_transposed and reverse.gather_indices.data_ptr() == forward.scatter_indices.data_ptr()
    cpp.validate_gather_scatter_default_topology(fine.data, coarse, reverse)
    cpp.validate_gather_scatter_default_topology(fine.data, coarse, cpp.gs_reverse_topology(reverse))
    big = _grid(fvdb, _random_batch(21, n=2000, extent=12, batches=2))
    for k, s_ in ((3, 1), (2, 2), ((3, 1, 2), (1, 2, 1))):
        target = big.data if s_ == 1 else cpp.conv_grid(big.data, oracle.normalize_3d(k), oracle.normalize_3d(s_))
        cpp.validate_gather_scatter_default_topology(big.data, target, cpp.gs_build_topology(big.data, target, oracle.normalize_3d(k), oracle.normalize_3d(s_)))
    dense = _grid(fvdb, [[(x, y, z) for x in range(2) for y in range(2) for z in range(2)]])
    topo = cpp.gs_build_topology(dense.data, dense.data, (1, 1, 1), (1, 1, 1))
    assert topo.total_pairs > 1
    cpp.validate_gather_scatter_default_topology(dense.data, dense.data, topo)
    with pytest.raises(RuntimeError, match="feature voxel count"):
        cpp.validate_gather_scatter_default_topology(dense.data, dense.data, topo, feature_total_voxels=topo.feature_total_voxels + 1)
    bad_offsets = topo.offsets.clone()
    bad_offsets[0] = 1
    with pytest.raises(RuntimeError, match="offsets must start at zero"):
        cpp.validate_gather_scatter_default_topology(dense.data, dense.data, topo, offsets=bad_offsets)
    with pytest.raises(RuntimeError, match="gather index out of range"):
        cpp.validate_gather_scatter_default_topology(dense.data, dense.data, topo, gather_indices=torch.full_like(topo.gather_indices, topo.feature_total_voxels))
    bad_edge = topo.gather_indices.clone()
    bad_edge[0] = (int(bad_edge[0]) + 1) % topo.feature_total_voxels
    with pytest.raises(RuntimeError, match="canonical fine/coarse geometry"):
        cpp.validate_gather_scatter_default_topology(dense.data, dense.data, topo, gather_indices=bad_edge)


def test_kernel_map_wide_neighbourhood_falls_back_to_tree_walk(fvdb):
    # stride 5, kernel 6: the probe box of one output leaf spans > 128 source leaves -> per-probe tree walk path
    fine = _grid(fvdb, [_random_batch(4, n=4000, extent=60, batches=1)[0]])
    coarse = fine.conv_grid(6, 5)
    topo = fvdb._fvdb_cpp.gs_build_topology(fine.data, coarse.data, [6] * 3, [5] * 3)
    want = oracle.dense_kernel_map(*_rows(fine), *_rows(coarse), 6, 5)
    assert np.array_equal(topo._out_map().cpu().numpy()[:, : coarse.total_voxels].T, want)


# ------------------------------------------------------------------ values and gradients


def test_golden_dense_values_and_gradients_fp64(fvdb, dense_golden):
    data, meta = dense_golden
    for case in meta:
        key, ks, st, transposed = case["key"], case["kernel_size"], case["stride"], case["transposed"]
        source = _grid(fvdb, [data[key + "_source"]])
        factory = fvdb.ConvolutionPlan.from_grid_batch_transposed if transposed else fvdb.ConvolutionPlan.from_grid_batch
        plan = factory(kernel_size=ks, stride=st, source_grid=source, acknowledge_incomplete_coverage=True)
        src_rows = {tuple(r): i for i, r in enumerate(data[key + "_source"].tolist())}
        tgt_rows = {tuple(r): i for i, r in enumerate(data[key + "_target"].tolist())}
        src_perm = [src_rows[tuple(r)] for r in source.ijk.jdata.cpu().tolist()]
        tgt_perm = [tgt_rows[tuple(r)] for r in plan.target_grid_batch.ijk.jdata.cpu().tolist()]
        assert sorted(tgt_perm) == list(range(len(tgt_rows)))  # generated target == reference support
        x = torch.from_numpy(data[key + "_features"])[src_perm].to(DEV).requires_grad_()
        w = torch.from_numpy(data[key + "_weights"]).to(DEV).requires_grad_()
        y = plan.execute(x, w)
        torch.testing.assert_close(y.detach().cpu(), torch.from_numpy(data[key + "_values"])[tgt_perm], rtol=1e-11, atol=1e-11)
        probe = torch.arange(1, y.numel() + 1, dtype=torch.float64).reshape(len(tgt_rows), -1)[tgt_perm].to(DEV)
        gx, gw = torch.autograd.grad((y * probe).sum(), (x, w))
        torch.testing.assert_close(gx.cpu(), torch.from_numpy(data[key + "_grad_features"])[src_perm], rtol=1e-11, atol=1e-11)
        torch.testing.assert_close(gw.cpu(), torch.from_numpy(data[key + "_grad_weights"]), rtol=1e-11, atol=1e-11)


def _oracle_run(plan_topology, x, w, dy):
    topo = oracle.Topology(
        plan_topology.gather_indices.cpu().numpy(), plan_topology.scatter_indices.cpu().numpy(), plan_topology.offsets.numpy(),
        plan_topology.feature_total_voxels, plan_topology.output_total_voxels, plan_topology.kernel_volume, plan_topology.total_pairs,
        tuple(plan_topology.kernel_size), tuple(plan_topology.stride), plan_topology.is_transposed)
    y = oracle.gs_conv(x.float().cpu(), w.float().cpu(), topo, accumulate_dtype=torch.float32)
    gx, gw = oracle.gs_conv_backward(dy.float().cpu(), x.float().cpu(), w.float().cpu(), topo, accumulate_dtype=torch.float32)
    return y, gx, gw


def _rel_err(got, want):
    return float((got.float().cpu() - want).norm() / want.norm().clamp_min(1e-30))


_VALUE_CASES = [
    (torch.float32, 32, 32, 3, 1, 1e-5), (torch.float32, 4, 16, 3, 1, 1e-5), (torch.float32, 16, 48, (3, 5, 1), (1, 2, 1), 1e-5),
    (torch.float32, 64, 64, 2, 2, 1e-5), (torch.float32, 128, 96, 3, 1, 1e-5), (torch.float32, 3, 5, 3, 1, 1e-5),
    (torch.bfloat16, 64, 64, 3, 1, 2e-2), (torch.bfloat16, 32, 32, 3, 1, 2e-2), (torch.bfloat16, 16, 16, 5, 1, 2e-2),
    (torch.bfloat16, 128, 128, 3, 1, 2e-2), (torch.bfloat16, 64, 128, 2, 2, 2e-2), (torch.bfloat16, 256, 256, 3, 1, 2e-2),
    (torch.bfloat16, 64, 32, 3, 2, 2e-2), (torch.float16, 64, 64, 3, 1, 2e-2), (torch.float16, 32, 64, 3, 1, 2e-2),
    (torch.bfloat16, 24, 40, 3, 1, 2e-2),
    # packed small-channel tensor-core path: Cin < 64 puts 64 / Cin taps in one reduction block
    (torch.bfloat16, 16, 32, 3, 1, 2e-2), (torch.bfloat16, 32, 16, 2, 2, 2e-2), (torch.bfloat16, 16, 64, (3, 5, 1), (1, 2, 1), 2e-2),
    (torch.float16, 16, 16, 3, 1, 2e-2), (torch.bfloat16, 128, 16, 3, 1, 2e-2), (torch.bfloat16, 32, 256, 3, 1, 2e-2),
]


@pytest.mark.parametrize("dtype,cin,cout,ks,st,tol", _VALUE_CASES)
@pytest.mark.parametrize("transposed", [False, True])
def test_values_and_gradients_match_oracle(fvdb, dtype, cin, cout, ks, st, tol, transposed):
    from fvdb.utils.synthetic import sphere_shell

    shell = sphere_shell(target=6000, domain=64, seed=3, device="cpu").numpy()
    source = _grid(fvdb, [shell, _random_batch(5, n=1500, extent=10, batches=1)[0]])
    factory = fvdb.ConvolutionPlan.from_grid_batch_transposed if transposed else fvdb.ConvolutionPlan.from_grid_batch
    same_topology = (not transposed) and oracle.normalize_3d(st) == (1, 1, 1)
    plan = factory(kernel_size=ks, stride=st, source_grid=source, target_grid=source if same_topology else None, acknowledge_incomplete_coverage=True)
    topo = plan._backend.topology
    gen = torch.Generator().manual_seed(42)
    k = oracle.normalize_3d(ks)
    x = torch.randn((source.total_voxels, cin), generator=gen).to(dtype).to(DEV).requires_grad_()
    bound = 1.0 / (cin * k[0] * k[1] * k[2]) ** 0.5
    w = ((torch.rand((cout, cin, *k), generator=gen) * 2 - 1) * bound).to(dtype).to(DEV).requires_grad_()
    dy = torch.randn((plan.target_grid_batch.total_voxels, cout), generator=gen).to(dtype).to(DEV)
    y = plan.execute(source.jagged_like(x), w)
    assert y.jdata.dtype == dtype and y.jdata.shape == (plan.target_grid_batch.total_voxels, cout)
    gx, gw = torch.autograd.grad(y.jdata, (x, w), dy)
    want_y, want_gx, want_gw = _oracle_run(topo, x.detach(), w.detach(), dy)
    assert _rel_err(y.jdata.detach(), want_y) <= tol
    assert _rel_err(gx, want_gx) <= tol
    assert _rel_err(gw, want_gw) <= tol * (4 if dtype == torch.float32 else 1)  # reference widens kernel-grad tolerance (convolution_utils.py:119-131)
    if dtype == torch.float32:
        # elementwise too, against the fp64-accumulated oracle, at the reference's own fp32 bars (forward / input gradient
        # rtol 1e-5, atol 1e-6; kernel gradient 5e-4 / 5e-4: fvdb/utils/tests/convolution_utils.py:115-136).  Those bars were
        # validated there for 1-8 feature channels and scale the gradient bars with sqrt(kernel volume / 27); rounding error grows
        # with the square root of the number of accumulated terms, so the absolute bar is 2e-6 * sqrt(terms / (8 * 27)) here (the CUDA-core
        # kernels, which serve the odd channel counts, accumulate sequentially in fp32)
        # (C1's own 32-channel 3^3 case, tests/test_gpu_at_size.py, passes the unscaled 1e-5 / 1e-6).
        topo_ref = oracle.Topology(topo.gather_indices.cpu().numpy(), topo.scatter_indices.cpu().numpy(), topo.offsets.numpy(), topo.feature_total_voxels,
                                   topo.output_total_voxels, topo.kernel_volume, topo.total_pairs, tuple(topo.kernel_size), tuple(topo.stride), topo.is_transposed)
        xd, wd, dyd = x.detach().double().cpu(), w.detach().double().cpu(), dy.double().cpu()
        true_y = oracle.gs_conv(xd, wd, topo_ref, accumulate_dtype=torch.float64)
        true_gx, true_gw = oracle.gs_conv_backward(dyd, xd, wd, topo_ref, accumulate_dtype=torch.float64)
        k3 = k[0] * k[1] * k[2]
        torch.testing.assert_close(y.jdata.detach().cpu().double(), true_y, rtol=1e-5, atol=2e-6 * max(1.0, (cin * k3 / 216.0) ** 0.5))
        torch.testing.assert_close(gx.cpu().double(), true_gx, rtol=1e-5, atol=2e-6 * max(1.0, (cout * k3 / 216.0) ** 0.5))
        torch.testing.assert_close(gw.cpu().double(), true_gw, rtol=5e-4 * max(1.0, (k3 / 27.0) ** 0.5), atol=5e-4 * max(1.0, (k3 / 27.0) ** 0.5))


@pytest.mark.parametrize("dtype,cin,cout,ks,st", [(torch.bfloat16, 16, 16, 5, 1), (torch.bfloat16, 32, 32, 3, 1), (torch.float16, 16, 32, 3, 1),
                                                   (torch.bfloat16, 32, 16, 3, 1), (torch.bfloat16, 32, 32, 2, 2), (torch.bfloat16, 16, 16, (3, 5, 1), (1, 2, 1)),
                                                   (torch.float16, 16, 16, 3, 2)])
@pytest.mark.parametrize("transposed", [False, True])
def test_fused_backward_matches_separate_kernels_and_oracle(fvdb, dtype, cin, cout, ks, st, transposed):
    """fvc_conv_backward_fused (narrow layers: dgrad and wgrad off one gather of grad_output) against the two separate
    kernels on the same inputs, and both against the oracle; large enough that every persistent CTA walks several tiles, with a
    far-away second grid so some tiles reach only a few taps."""
    from fvdb import _fvdb_cpp
    from fvdb.utils.synthetic import sphere_shell

    _FUSED_DEFAULT = _fvdb_cpp._fused_backward
    shell = sphere_shell(target=60000, domain=160, seed=11, device="cpu").numpy()
    source = _grid(fvdb, [shell, _random_batch(7, n=4000, extent=12, batches=1)[0] + 500])
    factory = fvdb.ConvolutionPlan.from_grid_batch_transposed if transposed else fvdb.ConvolutionPlan.from_grid_batch
    same_topology = (not transposed) and oracle.normalize_3d(st) == (1, 1, 1)
    plan = factory(kernel_size=ks, stride=st, source_grid=source, target_grid=source if same_topology else None, acknowledge_incomplete_coverage=True)
    topo = plan._backend.topology
    k = oracle.normalize_3d(ks)
    k3 = k[0] * k[1] * k[2]
    assert int(_fvdb_cpp.lib.fvc_conv_kernel_family(cin, cout, k3, _fvdb_cpp._DTYPE_CODE[dtype], 0, 2)) == 2
    gen = torch.Generator().manual_seed(17)
    x = torch.randn((source.total_voxels, cin), generator=gen).to(dtype).to(DEV)
    w = ((torch.rand((cout, cin, *k), generator=gen) * 2 - 1) / (cin * k3) ** 0.5).to(dtype).to(DEV)
    dy = torch.randn((plan.target_grid_batch.total_voxels, cout), generator=gen).to(dtype).to(DEV)
    backward = _fvdb_cpp.gs_conv_transpose_backward if transposed else _fvdb_cpp.gs_conv_backward
    _fvdb_cpp.set_fused_backward(True)
    gx_f, gw_f = backward(dy, x, w, topo)  # (the first call also builds the lazily derived input-stationary map)
    before = int(_fvdb_cpp.lib.fvc_launch_count())
    gx_f2, gw_f2 = backward(dy, x, w, topo)
    fused_launches = int(_fvdb_cpp.lib.fvc_launch_count()) - before
    assert torch.equal(gx_f, gx_f2) and torch.equal(gw_f, gw_f2)  # no atomics: run-to-run deterministic
    _fvdb_cpp.set_fused_backward(False)
    try:
        before = int(_fvdb_cpp.lib.fvc_launch_count())
        gx_s, gw_s = backward(dy, x, w, topo)
        separate_launches = int(_fvdb_cpp.lib.fvc_launch_count()) - before
    finally:
        _fvdb_cpp.set_fused_backward(_FUSED_DEFAULT)
    assert fused_launches < separate_launches  # the fused kernel really served the first call
    # same products, fp32 accumulation in a different order, one rounding to the half type at the end
    assert _rel_err(gx_f, gx_s.float().cpu()) <= 4e-3 and _rel_err(gw_f, gw_s.float().cpu()) <= 4e-3
    _, want_gx, want_gw = _oracle_run(topo, x, w, dy)
    assert _rel_err(gx_f, want_gx) <= 2e-2 and _rel_err(gw_f, want_gw) <= 2e-2
    torch.testing.assert_close(gx_f.float().cpu(), want_gx, rtol=2e-2, atol=2e-2 * float(want_gx.abs().max()))


@pytest.mark.parametrize("dtype,cin,cout,tol", [(torch.bfloat16, 64, 64, 2e-2), (torch.bfloat16, 128, 32, 2e-2), (torch.float16, 64, 128, 2e-2),
                                                 (torch.bfloat16, 32, 64, 2e-2), (torch.bfloat16, 16, 16, 2e-2), (torch.float32, 64, 64, 2e-5),
                                                 (torch.float32, 32, 128, 2e-5), (torch.bfloat16, 256, 256, 2e-2)])
def test_fused_block_epilogue_matches_separate_passes(fvdb, dtype, cin, cout, tol):
    # conv -> (+bias) -> BatchNorm-apply (scale, shift) -> (+residual) -> ReLU in the GEMM epilogue (fvdb/nn/modules.py:484-521,
    # simple_unet.py:233-243) against the same chain as separate fp32 passes over the oracle's convolution; and the per-block
    # column sums against a statistics pass over the stored output
    from fvdb import _norm
    from fvdb.utils.synthetic import sphere_shell

    cpp = fvdb._fvdb_cpp
    source = _grid(fvdb, [sphere_shell(target=9000, domain=64, seed=4, device="cpu").numpy(), _random_batch(8, n=700, extent=9, batches=1)[0]])
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, source, source)
    topo = plan._backend.topology
    n = source.total_voxels
    gen = torch.Generator().manual_seed(9)
    x = torch.randn((n, cin), generator=gen).to(dtype).to(DEV)
    w = ((torch.rand((cout, cin, 3, 3, 3), generator=gen) * 2 - 1) / (cin * 27) ** 0.5).to(dtype).to(DEV)
    bias = torch.randn(cout, generator=gen).to(dtype).to(DEV)
    scale = (torch.rand(cout, generator=gen) + 0.5).to(DEV)
    shift = torch.randn(cout, generator=gen).to(DEV)
    res = torch.randn((n, cout), generator=gen).to(dtype).to(DEV)
    conv, _, _ = _oracle_run(topo, x, w, torch.zeros((n, cout)))
    for use in ({"bias": bias}, {"scale": scale, "shift": shift}, {"relu": True}, {"residual": res}, {"residual": res, "relu": 2},
                {"bias": bias, "scale": scale, "shift": shift, "residual": res, "relu": 3}):
        want = conv.clone()
        if "bias" in use:
            want = want + bias.float().cpu()
        if "scale" in use:
            want = want * scale.cpu() + shift.cpu()
        if int(use.get("relu", 0)) & 1:  # the block's activation, before the skip connection joins
            want = torch.relu(want)
        if "residual" in use:
            want = want + res.float().cpu()
        if int(use.get("relu", 0)) & 2:  # ... and the ReLU after it (fvdb/nn/simple_unet.py:187-188)
            want = torch.relu(want)
        y, stats = cpp.gs_conv(x, w, topo, **use, want_stats=True)
        assert y.dtype == dtype and _rel_err(y, want) <= tol, (sorted(use), _rel_err(y, want))
        stored = y.double()
        part = stats.partial.double()
        assert stats.rows == n and part.shape[0] == (n + stats.rows_per_block - 1) // stats.rows_per_block
        torch.testing.assert_close(part[:, 0].sum(0), stored.sum(0), rtol=1e-4, atol=1e-3)
        torch.testing.assert_close(part[:, 1].sum(0), stored.square().sum(0), rtol=1e-4, atol=1e-3)
        mean, var = _norm.stats_from_conv_partials(stats)
        torch.testing.assert_close(mean.double(), stored.mean(0), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(var.double(), stored.var(0, unbiased=False), rtol=1e-3, atol=1e-6)
        again, stats2 = cpp.gs_conv(x, w, topo, **use, want_stats=True)
        assert torch.equal(again, y) and torch.equal(stats2.partial, stats.partial)  # deterministic, no atomics


@pytest.mark.parametrize("dtype,cin,cout,tol", [(torch.bfloat16, 64, 64, 2e-2), (torch.float32, 32, 64, 1e-4), (torch.bfloat16, 32, 32, 2e-2)])
def test_conv_bn_act_block_matches_the_separate_modules(fvdb, dtype, cin, cout, tol):
    # fvdb.nn.conv_bn_act == BatchNorm(activation)(SparseConv3d(x)) in training (values, running statistics, every gradient) and
    # == the eval-mode chain (+ residual) in inference, where the whole block is one kernel
    import copy

    from fvdb.utils.synthetic import sphere_shell

    source = _grid(fvdb, [sphere_shell(target=12_000, domain=64, seed=6, device="cpu").numpy(), _random_batch(10, n=900, extent=9, batches=1)[0]])
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, source, source)
    torch.manual_seed(5)
    conv = fvdb.nn.SparseConv3d(cin, cout, 3).to(DEV).to(dtype)
    norm = fvdb.nn.BatchNorm(cout, activation="relu").to(DEV).to(dtype)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5), norm.bias.uniform_(-0.5, 0.5)
    conv2, norm2 = copy.deepcopy(conv), copy.deepcopy(norm)
    x = torch.randn((source.total_voxels, cin), device=DEV).to(dtype)
    dy = torch.randn((source.total_voxels, cout), device=DEV).to(dtype)
    xa, xb = x.clone().requires_grad_(), x.clone().requires_grad_()
    out = fvdb.nn.conv_bn_act(conv, norm, source.jagged_like(xa), plan).jdata
    want = norm2(conv2(source.jagged_like(xb), plan)).jdata
    assert _rel_err(out, want.float().cpu()) <= tol
    out.backward(dy), want.backward(dy)
    assert _rel_err(xa.grad, xb.grad.float().cpu()) <= tol
    for p, q in zip([conv.weight] + list(norm.parameters()), [conv2.weight] + list(norm2.parameters())):
        torch.testing.assert_close(p.grad.float(), q.grad.float(), rtol=5 * tol, atol=5 * tol * float(q.grad.float().abs().max()))
    # a bias in front of a training-mode BatchNorm cancels: the fused block reports its gradient as exactly zero, the separate
    # modules compute rounding noise around zero
    assert int(torch.count_nonzero(conv.bias.grad)) == 0
    assert float(conv2.bias.grad.float().abs().max()) <= 2e-2 * float(dy.float().abs().sum(0).max())
    torch.testing.assert_close(norm.running_mean.float(), norm2.running_mean.float(), rtol=1e-2, atol=1e-3)
    torch.testing.assert_close(norm.running_var.float(), norm2.running_var.float(), rtol=1e-2, atol=1e-3)
    conv.eval(), norm.eval(), conv2.eval(), norm2.eval()
    res = source.jagged_like(torch.randn((source.total_voxels, cout), device=DEV).to(dtype))
    with torch.no_grad():
        launches = fvdb._lib.launch_count()
        fused = fvdb.nn.conv_bn_act(conv, norm, source.jagged_like(x), plan, residual=res, final_relu=True).jdata
        assert fvdb._lib.launch_count() - launches <= 3  # weight image (+ fp32 row split) + ONE convolution kernel for the whole block
        # the tail of the reference's residual block: relu(relu(norm(conv(x))) + skip)  (fvdb/nn/simple_unet.py:182-188)
        chain = torch.relu(torch.relu(torch.nn.functional.batch_norm(conv2(source.jagged_like(x), plan).jdata.float(), norm2.running_mean.float(), norm2.running_var.float(),
                                                                     norm2.weight.float(), norm2.bias.float(), False, 0.0, norm2.eps)) + res.jdata.float())
    assert _rel_err(fused, chain.cpu()) <= tol


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 13, 14])
@pytest.mark.parametrize("channels", [64, 128])
def test_forward_pipeline_variants_agree_with_oracle(fvdb, variant, channels):
    # the bench knob's pipeline shapes (warp-per-unit producers with 2 / 3 / 4 warps, 1..3 CTAs per SM, and the round-1 ring kernel)
    from fvdb.utils.synthetic import sphere_shell

    cpp = fvdb._fvdb_cpp
    source = _grid(fvdb, [sphere_shell(target=20_000, domain=96, seed=5, device="cpu").numpy(), _random_batch(9, n=1900, extent=11, batches=1)[0]])
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, source, source)
    topo = plan._backend.topology
    gen = torch.Generator().manual_seed(12)
    x = torch.randn((source.total_voxels, channels), generator=gen).bfloat16().to(DEV)
    w = ((torch.rand((channels, channels, 3, 3, 3), generator=gen) * 2 - 1) / (channels * 27) ** 0.5).bfloat16().to(DEV)
    dy = torch.randn((source.total_voxels, channels), generator=gen).bfloat16().to(DEV)
    want_y, want_gx, _ = _oracle_run(topo, x, w, dy)
    try:
        cpp.set_kernel_variant(variant)
        y = cpp.gs_conv(x, w, topo)
        gx, _ = cpp.gs_conv_backward(dy, x, w, topo)
    finally:
        cpp.set_kernel_variant(0)
    default = cpp.gs_conv(x, w, topo)
    assert _rel_err(y, want_y) <= 2e-2 and _rel_err(gx, want_gx) <= 2e-2
    assert torch.equal(y, default)  # same per-row MMA order in every shape


@pytest.mark.parametrize("dtype,cin,cout,ks,st", [(torch.bfloat16, 16, 16, 5, 1), (torch.bfloat16, 32, 32, 3, 1), (torch.float16, 16, 32, 3, 1),
                                                   (torch.bfloat16, 32, 16, 2, 2), (torch.bfloat16, 16, 16, (3, 5, 1), (1, 2, 1))])
def test_tensor_memory_executor_matches_shared_memory_kernels(fvdb, dtype, cin, cout, ks, st):
    """Variant 12: narrow layers with the gathered operand written to tensor memory by tcgen05.st (conv_tc_ts.cu) -- forward, dgrad,
    the fused block epilogue and its statistics -- against the default kernels and the oracle."""
    from fvdb.utils.synthetic import sphere_shell

    cpp = fvdb._fvdb_cpp
    source = _grid(fvdb, [sphere_shell(target=60_000, domain=160, seed=11, device="cpu").numpy(), _random_batch(7, n=4000, extent=12, batches=1)[0] + 500])
    same = oracle.normalize_3d(st) == (1, 1, 1)
    plan = fvdb.ConvolutionPlan.from_grid_batch(ks, st, source, source if same else None, acknowledge_incomplete_coverage=True)
    topo = plan._backend.topology
    k = oracle.normalize_3d(ks)
    gen = torch.Generator().manual_seed(23)
    n_out = plan.target_grid_batch.total_voxels
    x = torch.randn((source.total_voxels, cin), generator=gen).to(dtype).to(DEV)
    w = ((torch.rand((cout, cin, *k), generator=gen) * 2 - 1) / (cin * k[0] * k[1] * k[2]) ** 0.5).to(dtype).to(DEV)
    dy = torch.randn((n_out, cout), generator=gen).to(dtype).to(DEV)
    bias = torch.randn(cout, generator=gen).to(dtype).to(DEV)
    scale, shift = torch.rand(cout, generator=gen).to(DEV) + 0.5, torch.randn(cout, generator=gen).to(DEV)
    res = torch.randn((n_out, cout), generator=gen).to(dtype).to(DEV)

    def run():
        y = cpp.gs_conv(x, w, topo, bias)
        gx, _ = cpp.gs_conv_backward(dy, x, w, topo)
        z, stats = cpp.gs_conv(x, w, topo, bias, scale=scale, shift=shift, residual=res, relu=3, want_stats=True)
        sums = stats.partial.double().sum(0)  # [2, cout]: column sums / sums of squares of the stored values
        return y, gx, z, sums

    want = run()
    try:
        cpp.set_kernel_variant(12)
        got = run()
        again = run()
    finally:
        cpp.set_kernel_variant(0)
    want_y, want_gx, _ = _oracle_run(topo, x, w, dy)
    assert _rel_err(got[0].float() - bias.float(), want_y) <= 2e-2 and _rel_err(got[1], want_gx) <= 2e-2
    for a, b, c in zip(got[:3], want[:3], again[:3]):
        assert _rel_err(a, b.float().cpu()) <= 4e-3  # same products, fp32 sums in another order, one rounding at the end
        assert torch.equal(a, c)  # deterministic
    torch.testing.assert_close(got[3], want[3], rtol=2e-3, atol=2e-3 * float(want[3].abs().max()))
    z = got[2].double()
    torch.testing.assert_close(got[3], torch.stack([z.sum(0), (z * z).sum(0)]), rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("transposed", [False, True])
@pytest.mark.parametrize("stride", [1, (2, 2, 2)])
def test_gradcheck_fp64(fvdb, transposed, stride):
    # torch.autograd.gradcheck of the autograd glue in fp64, forward and transposed, stride 1 and (2, 2, 2), at the reference's
    # own tolerances (/root/reference/tests/unit/test_conv_default.py:775-850)
    from fvdb.convolution_plan import _GatherScatterConvFn

    rng = np.random.default_rng(17)
    coords = np.unique(rng.integers(-3, 4, size=(40, 3)), axis=0)
    grid = _grid(fvdb, [coords])
    ks = (3, 3, 3)
    if transposed:
        dst = grid.conv_transpose_grid(kernel_size=ks, stride=stride)
        plan = fvdb.ConvolutionPlan.from_grid_batch_transposed(kernel_size=ks, stride=stride, source_grid=grid, target_grid=dst)
    else:
        dst = grid.conv_grid(kernel_size=ks, stride=stride)
        plan = fvdb.ConvolutionPlan.from_grid_batch(kernel_size=ks, stride=stride, source_grid=grid, target_grid=dst)
    topo = plan._backend.topology
    gen = torch.Generator().manual_seed(3)
    features = torch.randn((grid.total_voxels, 2), generator=gen, dtype=torch.float64).to(DEV).requires_grad_()
    weights = torch.randn((3, 2, *ks), generator=gen, dtype=torch.float64).to(DEV).requires_grad_()
    assert torch.autograd.gradcheck(lambda f, w: _GatherScatterConvFn.apply(f, w, None, topo, transposed), (features, weights), eps=1e-6, atol=1e-4, rtol=1e-3)


def test_forced_cuda_core_path_for_half(fvdb):
    cpp = fvdb._fvdb_cpp
    source = _grid(fvdb, [_random_batch(6, n=2500, extent=9, batches=1)[0]])
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, source, source)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn((source.total_voxels, 64), generator=gen).bfloat16().to(DEV)
    w = (torch.randn((64, 64, 3, 3, 3), generator=gen) * 0.03).bfloat16().to(DEV)
    dy = torch.randn((source.total_voxels, 64), generator=gen).bfloat16().to(DEV)
    try:
        cpp.set_conv_path("simt")
        y = cpp.gs_conv(x, w, plan._backend.topology)
        gx, gw = cpp.gs_conv_backward(dy, x, w, plan._backend.topology)
    finally:
        cpp.set_conv_path("auto")
    want_y, want_gx, want_gw = _oracle_run(plan._backend.topology, x, w, dy)
    assert _rel_err(y, want_y) <= 2e-2 and _rel_err(gx, want_gx) <= 2e-2 and _rel_err(gw, want_gw) <= 2e-2


def test_all_ones_equals_rulebook_degree_and_adjoint(fvdb):
    # size-independent properties on a larger grid: all-ones => degree (test_conv_semantics_integration.py:149-168);
    # weighted adjoint <y, d> == <x, L^T d> through from_plan_transposed (:903-947)
    from fvdb.utils.synthetic import indoor_room

    source = _grid(fvdb, [indoor_room(target=60_000, seed=s, device="cpu").numpy() for s in (0, 1)])
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 2, source, acknowledge_incomplete_coverage=True)
    ones = plan.execute(source.jagged_like(torch.ones((source.total_voxels, 1), dtype=torch.float64, device=DEV)), torch.ones((1, 1, 3, 3, 3), dtype=torch.float64, device=DEV))
    degree = torch.bincount(plan._backend.topology.scatter_indices, minlength=plan.target_grid_batch.total_voxels)
    assert torch.equal(ones.jdata[:, 0], degree.double()) and int(degree.min()) >= 1
    report = plan.coverage_report
    assert report.output_zero_count == 0 and report.input_row_count == source.total_voxels
    gen = torch.Generator().manual_seed(50)
    x = torch.randn((source.total_voxels, 8), generator=gen, dtype=torch.float64).to(DEV)
    w = torch.randn((16, 8, 3, 3, 3), generator=gen, dtype=torch.float64).to(DEV)
    d = torch.randn((plan.target_grid_batch.total_voxels, 16), generator=gen, dtype=torch.float64).to(DEV)
    y = plan.execute(source.jagged_like(x), w).jdata
    adj = fvdb.ConvolutionPlan.from_plan_transposed(plan)
    assert adj.topology_provenance is fvdb.ConvolutionTopologyProvenance.EXACT_TRANSPOSE
    lt_d = adj.execute(plan.target_grid_batch.jagged_like(d), w.transpose(0, 1).contiguous()).jdata
    torch.testing.assert_close((y * d).sum(), (x * lt_d).sum(), rtol=1e-12, atol=1e-8)


def test_flip_identity_and_determinism(fvdb):
    source = _grid(fvdb, [_random_batch(7, n=4000, extent=12, batches=1)[0]])
    fwd = fvdb.ConvolutionPlan.from_grid_batch(3, 1, source, source)
    tr = fvdb.ConvolutionPlan.from_grid_batch_transposed(3, 1, source, source)
    gen = torch.Generator().manual_seed(70)
    x = torch.randn((source.total_voxels, 8), generator=gen, dtype=torch.float64).to(DEV)
    w = torch.randn((8, 8, 3, 3, 3), generator=gen, dtype=torch.float64).to(DEV)
    torch.testing.assert_close(tr.execute(x, w), fwd.execute(x, w.flip(2, 3, 4)), rtol=1e-12, atol=1e-12)
    xb, wb = x.bfloat16(), (w * 0.05).bfloat16()
    assert torch.equal(fwd.execute(xb, wb), fwd.execute(xb, wb))  # no atomics on the output-stationary path
    dy = torch.randn_like(xb)
    g1 = fvdb._fvdb_cpp.gs_conv_backward(dy, xb, wb, fwd._backend.topology)
    g2 = fvdb._fvdb_cpp.gs_conv_backward(dy, xb, wb, fwd._backend.topology)
    assert torch.equal(g1[0], g2[0]) and torch.equal(g1[1], g2[1])


# ------------------------------------------------------------------ plan / module behaviour


def test_plan_api_contract(fvdb):
    Plan = fvdb.ConvolutionPlan
    fine = _grid(fvdb, [[(0, 0, 0)]], voxel_sizes=(0.5, 1.0, 2.0), origins=(3.0, -2.0, 7.0))
    plan = Plan.from_grid_batch(kernel_size=(4, 3, 2), stride=1, source_grid=fine)
    assert plan.geometry.kernel_size == [4, 3, 2] and plan.geometry.padding_before == [1, 1, 0] and plan.geometry.padding_after == [2, 1, 1]
    assert plan.geometry.kernel_volume == 24 and plan.geometry.semantics_version == 1 and plan.geometry.phase_policy == "torch_same_phase"
    assert plan.phase_policy is fvdb.ConvolutionPhasePolicy.TORCH_SAME_PHASE and plan.transform_compatibility.compatible
    strided = Plan.from_grid_batch(kernel_size=3, stride=2, source_grid=fine)
    torch.testing.assert_close(strided.target_grid_batch.voxel_sizes.cpu(), torch.tensor([[1.0, 2.0, 4.0]]))
    torch.testing.assert_close(strided.target_grid_batch.origins.cpu(), fine.origins.cpu())
    unit = _grid(fvdb, [[(0, 0, 0)]])
    with pytest.raises(ValueError, match="nonzero integer.*a=0"):
        Plan.from_grid_batch(3, 2, unit, _grid(fvdb, [[(0, 0, 0)]], voxel_sizes=2.0, origins=(1.0, 0.0, 0.0)))
    with pytest.raises(ValueError, match="fractional.*a=0"):
        Plan.from_grid_batch(3, 2, unit, _grid(fvdb, [[(0, 0, 0)]], voxel_sizes=2.0, origins=(0.5, 0.0, 0.0)))
    with pytest.raises(ValueError, match="voxel size"):
        Plan.from_grid_batch(3, 2, unit, unit)
    with pytest.raises(ValueError, match="same batch size"):
        Plan.from_grid_batch(3, 1, unit, _grid(fvdb, [[(0, 0, 0)], [(1, 1, 1)]]))
    with pytest.raises(ValueError, match="COMPLETE.*target_grid=None"):
        Plan.from_grid_batch(3, 1, unit, unit, topology_policy=fvdb.ConvolutionTopologyPolicy.COMPLETE)
    with pytest.raises(ValueError, match="RESTRICTED.*explicit target_grid"):
        Plan.from_grid_batch(3, 1, unit, topology_policy=fvdb.ConvolutionTopologyPolicy.RESTRICTED)
    with pytest.raises(ValueError, match="dense convolution backend is disabled"):
        Plan.from_grid_batch(3, 1, unit, expert_config={"backend": "dense"})
    with pytest.raises(ValueError, match="uniform kernel sizes 3, 5, 7"):
        Plan.from_grid_batch(4, 1, unit, expert_config={"backend": "pred_gather_igemm"})
    with pytest.raises(ValueError, match="channel counts divisible by 32"):
        Plan.from_grid_batch(3, 1, unit, expert_config={"backend": "pred_gather_igemm"}, channel_pairs=((8, 32),))
    with pytest.raises(ValueError, match="does not support transposed convolution"):
        Plan.from_grid_batch_transposed(3, 1, unit, expert_config={"backend": "pred_gather_igemm"})
    far = _grid(fvdb, [[(100, 100, 100)]])
    with pytest.raises(ValueError, match="zero-degree output"):
        Plan.from_grid_batch(3, 1, unit, far, strict_output_coverage=True)
    assert Plan.from_grid_batch(3, 1, unit, far).execute(torch.ones(1, 2, device=DEV), torch.ones(3, 2, 3, 3, 3, device=DEV)).abs().sum() == 0
    with pytest.warns(fvdb.ConvolutionCoverageWarning, match="uncovered stride residues"):
        Plan.from_grid_batch(1, 2, _grid(fvdb, [[(0, 0, 0), (1, 1, 1)]]))
    # K = S = 1 on the same grid object is a matmul; equal-looking distinct grids go through the map
    ident = Plan.from_grid_batch(1, 1, unit)
    assert type(ident._backend).__name__ == "_MatmulBackend" and unit.conv_grid(1, 1) is unit
    assert type(Plan.from_grid_batch(1, 1, unit, _grid(fvdb, [[(0, 0, 0)]]))._backend).__name__ == "_GatherScatterBackend"
    x = torch.randn(1, 4, device=DEV)
    w2 = torch.randn(6, 4, device=DEV)
    torch.testing.assert_close(ident.execute(x, w2), x @ w2.T)
    with pytest.raises(ValueError, match="batch size of 1"):
        Plan.from_grid_batch(3, 1, _grid(fvdb, [[(0, 0, 0)], [(1, 1, 1)]])).execute(torch.ones(2, 2, device=DEV), torch.ones(2, 2, 3, 3, 3, device=DEV))
    with pytest.raises(ValueError, match="not supported"):
        Plan.from_grid_batch(3, 1, unit, channel_pairs=((2, 4),)).execute(torch.ones(1, 2, device=DEV), torch.ones(3, 2, 3, 3, 3, device=DEV))
    topo = Plan.from_grid_batch(3, 1, unit)._backend.topology
    with pytest.raises(RuntimeError, match="requires topology with direction=Transposed"):
        fvdb._fvdb_cpp.gs_conv_transpose(torch.ones(1, 2, device=DEV), torch.ones(3, 2, 3, 3, 3, device=DEV), topo)
    with pytest.raises(RuntimeError, match="features.size\\(0\\)"):
        fvdb._fvdb_cpp.gs_conv(torch.ones(5, 2, device=DEV), torch.ones(3, 2, 3, 3, 3, device=DEV), topo)
    with pytest.raises(RuntimeError, match="kernel_size"):
        fvdb._fvdb_cpp.gs_conv(torch.ones(1, 2, device=DEV), torch.ones(3, 2, 5, 3, 3, device=DEV), topo)
    # dtype promotion table (GatherScatterDefaultConvTest.cu:1712-1786)
    assert fvdb._fvdb_cpp.gs_conv(torch.ones(1, 2, device=DEV, dtype=torch.bfloat16), torch.ones(3, 2, 3, 3, 3, device=DEV), topo).dtype == torch.float32
    assert fvdb._fvdb_cpp.gs_conv(torch.ones(1, 2, device=DEV), torch.ones(3, 2, 3, 3, 3, device=DEV, dtype=torch.float64), topo).dtype == torch.float64


def test_nn_modules_train_step(fvdb):
    source = _grid(fvdb, [_random_batch(8, n=2000, extent=10, batches=1)[0], _random_batch(9, n=1000, extent=8, batches=1)[0]])
    down = fvdb.ConvolutionPlan.from_grid_batch(2, 2, source)
    same = fvdb.ConvolutionPlan.from_grid_batch(3, 1, source, source)
    up = fvdb.ConvolutionPlan.from_plan_transposed(down)
    conv = fvdb.nn.SparseConv3d(8, 16, 3).to(DEV)
    pool = fvdb.nn.SparseConv3d(16, 32, 2, 2, bias=False).to(DEV)
    unpool = fvdb.nn.SparseConvTranspose3d(32, 8, 2, 2).to(DEV)
    assert conv.weight.shape == (16, 8, 3, 3, 3) and not conv.weight.is_contiguous()
    x = source.jagged_like(torch.randn(source.total_voxels, 8, device=DEV))
    out = unpool(pool(conv(x, same), down), up)
    assert out.jdata.shape == (source.total_voxels, 8) and torch.isfinite(out.jdata).all()
    out.jdata.square().mean().backward()
    for p in list(conv.parameters()) + list(pool.parameters()) + list(unpool.parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().sum() > 0
    with pytest.raises(ValueError, match="mismatched"):
        conv(x, down)
    # fused bias == separate bias add (reference modules.py:370-371), forward and gradient, on both kernel families
    for module, dtype in ((conv, torch.float32), (fvdb.nn.SparseConv3d(64, 64, 3).to(DEV).bfloat16(), torch.bfloat16)):
        xin = source.jagged_like(torch.randn(source.total_voxels, module.in_channels, device=DEV).to(dtype))
        module.zero_grad()
        fused = module(xin, same).jdata
        fused.float().square().mean().backward()
        fused_grad = module.bias.grad.clone()
        plain = same.execute(xin, module.weight).jdata.float() + module.bias.float()
        torch.testing.assert_close(fused.float(), plain, rtol=2e-2 if dtype == torch.bfloat16 else 1e-5, atol=2e-2 if dtype == torch.bfloat16 else 1e-5)
        want_grad = torch.autograd.grad(plain.square().mean(), module.bias)[0].float()
        torch.testing.assert_close(fused_grad.float(), want_grad, rtol=5e-2, atol=1e-3)
    # strided (non-contiguous) nn weights give the same result as a contiguous copy
    torch.testing.assert_close(same.execute(x, conv.weight).jdata, same.execute(x, conv.weight.detach().contiguous()).jdata)


@pytest.mark.parametrize("cin,cout", [(64, 64), (128, 128), (64, 32), (256, 64), (32, 32), (16, 16), (16, 128)])
def test_tensor_core_identity_map_is_plain_gemm(fvdb, cin, cout):
    # K^3 = 1 on two equal-looking but distinct grids: the kernel map is the identity, so the tcgen05 kernel
    # must reproduce x @ W^T -- isolates UMMA descriptors / swizzle / TMEM epilogue from the gather logic.
    cpp = fvdb._fvdb_cpp
    coords = _random_batch(11, n=700, extent=6, batches=1, dup=False)[0]
    a, b = _grid(fvdb, [coords]), _grid(fvdb, [coords])
    topo = cpp.gs_build_topology(a.data, b.data, [1, 1, 1], [1, 1, 1])
    assert topo.total_pairs == a.total_voxels
    gen = torch.Generator().manual_seed(3)
    x = torch.randn((a.total_voxels, cin), generator=gen).bfloat16().to(DEV)
    w = (torch.randn((cout, cin, 1, 1, 1), generator=gen) / cin**0.5).bfloat16().to(DEV)
    try:
        cpp.set_conv_path("tc")
        y = cpp.gs_conv(x, w, topo)
    finally:
        cpp.set_conv_path("auto")
    want = x.float() @ w[:, :, 0, 0, 0].float().T
    assert _rel_err(y, want.cpu()) <= 1e-2
    torch.testing.assert_close(y.float(), want, rtol=2e-2, atol=2e-2)

@pytest.mark.parametrize("cin,cout", [(64, 64), (32, 32), (16, 16), (128, 128), (16, 64), (256, 32)])
def test_fp32_tensor_core_split_identity_map_is_fp32_gemm(fvdb, cin, cout):
    # fp32 on the tensor pipe = three-way bf16 split (six exact products, two TMEM accumulators): on an identity map the
    # result must be the fp32 GEMM to ~1e-6, far inside the reference's 1e-5 bar and ~1000x tighter than one bf16 pass.
    cpp = fvdb._fvdb_cpp
    coords = _random_batch(12, n=900, extent=7, batches=1, dup=False)[0]
    a, b = _grid(fvdb, [coords]), _grid(fvdb, [coords])
    topo = cpp.gs_build_topology(a.data, b.data, [1, 1, 1], [1, 1, 1])
    gen = torch.Generator().manual_seed(4)
    x = torch.randn((a.total_voxels, cin), generator=gen).to(DEV)
    w = (torch.randn((cout, cin, 1, 1, 1), generator=gen) / cin**0.5).to(DEV)
    bias = torch.randn(cout, generator=gen).to(DEV)
    try:
        cpp.set_conv_path("tc")
        y = cpp.gs_conv(x, w, topo)
        yb = cpp.gs_conv(x, w, topo, bias)
    finally:
        cpp.set_conv_path("auto")
    want = (x.double() @ w[:, :, 0, 0, 0].double().T).float().cpu()
    assert y.dtype == torch.float32 and _rel_err(y, want) <= 2e-6
    torch.testing.assert_close(yb.cpu(), want + bias.cpu(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("positive", [False, True])
def test_fp32_tensor_core_split_matches_cuda_core_path(fvdb, positive):
    # Same plan through both kernel families; same-sign data is the worst case for the tensor pipe's truncating
    # accumulation (every step loses up to one ulp in the same direction).
    from fvdb.utils.synthetic import sphere_shell

    cpp = fvdb._fvdb_cpp
    grid = _grid(fvdb, [sphere_shell(target=9000, domain=64, seed=7, device="cpu").numpy(), _random_batch(8, n=3000, extent=9, batches=1)[0]])
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid, grid)
    topo = plan._backend.topology
    gen = torch.Generator().manual_seed(9)
    x = torch.randn((grid.total_voxels, 64), generator=gen)
    w = torch.randn((64, 64, 3, 3, 3), generator=gen) / 41.0
    dy = torch.randn((grid.total_voxels, 64), generator=gen)
    if positive:
        x, w, dy = x.abs(), w.abs(), dy.abs()
    x, w, dy = x.to(DEV), w.to(DEV), dy.to(DEV)
    got = {}
    for path in ("simt", "auto"):
        try:
            cpp.set_conv_path(path)
            got[path] = (cpp.gs_conv(x, w, topo), *cpp.gs_conv_backward(dy, x, w, topo))
        finally:
            cpp.set_conv_path("auto")
    for a, b, tol in zip(got["simt"], got["auto"], (5e-6, 5e-6, 1e-5)):  # y, grad_features, grad_weights (longer reduction)
        assert _rel_err(b, a.cpu()) <= tol
    want_y, want_gx, want_gw = _oracle_run(topo, x, w, dy)
    assert _rel_err(got["auto"][0], want_y) <= 1e-5 and _rel_err(got["auto"][1], want_gx) <= 1e-5


def test_pred_gather_igemm_backend_admission_and_values(fvdb):
    # reference: backend='pred_gather_igemm' (forward-only SM80 TF32 kernel, tests/unit/test_conv_pred_gather_igemm.py);
    # here the same engine serves it: admission rules kept, values equal the default backend's, backward works.
    coords = _random_batch(13, n=5000, extent=14, batches=1, dup=False)[0]
    grid = _grid(fvdb, [coords])
    igemm = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid, grid, expert_config={"backend": "pred_gather_igemm"}, channel_pairs=((64, 64),))
    default = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid, grid)
    assert type(igemm._backend).__name__ == "_PredGatherIGemmBackend" and igemm.valid_usage(64, 64, 3, 1, False) and not igemm.valid_usage(48, 64, 3, 1, False)
    gen = torch.Generator().manual_seed(8888)
    for dtype, tol in ((torch.float32, 1e-5), (torch.bfloat16, 2e-2)):
        x = torch.randn((grid.total_voxels, 64), generator=gen).to(dtype).to(DEV).requires_grad_()
        w = (torch.randn((64, 64, 3, 3, 3), generator=gen) / 41.0).to(dtype).to(DEV).requires_grad_()
        y = igemm.execute(x, w)
        torch.testing.assert_close(y, default.execute(x, w), rtol=0, atol=0)  # same kernels, deterministic
        want = _oracle_run(default._backend.topology, x.detach(), w.detach(), torch.ones_like(y))[0]
        assert _rel_err(y.detach(), want) <= tol
        gx, gw = torch.autograd.grad(y.float().sum(), (x, w))
        assert torch.isfinite(gx).all() and torch.isfinite(gw).all()
    direct = fvdb._fvdb_cpp.pred_gather_igemm_conv(x.detach(), w.detach(), grid.data, grid.data, 3, 1)
    torch.testing.assert_close(direct, default.execute(x.detach(), w.detach()), rtol=0, atol=0)
    any_pairs = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid, grid, expert_config={"backend": "pred_gather_igemm"})
    with pytest.raises(ValueError, match="channel counts divisible by 32"):
        any_pairs.execute(torch.ones(grid.total_voxels, 48, device=DEV), torch.ones(64, 48, 3, 3, 3, device=DEV))
    with pytest.raises(ValueError, match="only batch size 1"):
        fvdb.ConvolutionPlan.from_grid_batch(3, 1, _grid(fvdb, [coords, coords]), expert_config={"backend": "pred_gather_igemm"})


def test_host_pipelined_conv_matches_plain_execution(fvdb):
    from fvdb.streaming import HostPipelinedConv

    coords = [_random_batch(21 + i, n=6000, extent=16, batches=1, dup=False)[0] for i in range(3)]
    grid = _grid(fvdb, coords)
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid, grid)
    n = grid.total_voxels
    gen = torch.Generator().manual_seed(5)
    x_host = torch.randn((n, 64), generator=gen).bfloat16().pin_memory()
    dy_host = torch.randn((n, 64), generator=gen).bfloat16().pin_memory()
    w = (torch.randn((64, 64, 3, 3, 3), generator=gen) / 41.0).bfloat16().to(DEV)
    y_host, gx_host = torch.empty_like(dy_host).pin_memory(), torch.empty_like(x_host).pin_memory()
    gw_host = torch.empty(tuple(w.shape), dtype=torch.bfloat16).pin_memory()
    pipe = HostPipelinedConv(plan, num_chunks=5)
    assert len(pipe.bounds) >= 4 and all(r0 % 128 == 0 for r0, _ in pipe.bounds) and pipe.bounds[-1][1] == n
    pipe.forward_backward(x_host, dy_host, w, y_host, gx_host, gw_host)
    torch.cuda.synchronize()
    topo = plan._backend.topology
    y = fvdb._fvdb_cpp.gs_conv(x_host.to(DEV), w, topo)
    gx, gw = fvdb._fvdb_cpp.gs_conv_backward(dy_host.to(DEV), x_host.to(DEV), w, topo)
    assert torch.equal(y_host, y.cpu()) and torch.equal(gx_host, gx.cpu())  # same kernels on row sub-ranges: bit-identical
    assert _rel_err(gw_host, gw.float().cpu()) <= 1e-2  # chunk partials are summed in fp32, then rounded once


@pytest.mark.parametrize("dtype,channels,relu,tol", [(torch.float32, 32, False, 2e-5), (torch.float32, 64, True, 2e-5), (torch.float32, 24, True, 2e-5),
                                                     (torch.bfloat16, 64, True, 2e-2), (torch.bfloat16, 256, False, 2e-2), (torch.float16, 40, True, 5e-3)])
def test_batch_norm_matches_torch_batch_norm1d(fvdb, dtype, channels, relu, tol):
    # fvdb.nn.BatchNorm == torch.nn.BatchNorm1d over jdata (reference modules.py:484-521) [+ ReLU]: outputs, all three
    # gradients and the running statistics, in training and in eval mode.
    gen = torch.Generator().manual_seed(31)
    coords = [_random_batch(40 + i, n=5000, extent=14, batches=1, dup=False)[0] for i in range(2)]
    grid = _grid(fvdb, coords)
    n = grid.total_voxels
    x = (torch.randn((n, channels), generator=gen) * 1.7 + 0.6).to(dtype).to(DEV).requires_grad_()
    dy = torch.randn((n, channels), generator=gen).to(dtype).to(DEV)
    ours = fvdb.nn.BatchNorm(channels, activation="relu" if relu else None).to(DEV)
    ref = torch.nn.BatchNorm1d(channels).to(DEV)
    with torch.no_grad():
        ours.weight.copy_(torch.rand(channels, generator=gen) + 0.5)
        ours.bias.copy_(torch.randn(channels, generator=gen) * 0.3)
        ref.weight.copy_(ours.weight)
        ref.bias.copy_(ours.bias)
    xr = x.detach().float().requires_grad_()
    for step in range(2):  # two steps: running statistics accumulate
        y = ours(grid.jagged_like(x), grid).jdata
        yr = ref(xr)
        yr = torch.relu(yr) if relu else yr
    assert y.dtype == dtype and _rel_err(y.detach(), yr.detach().cpu()) <= tol
    gx, gw, gb = torch.autograd.grad(y, (x, ours.weight, ours.bias), dy)
    gxr, gwr, gbr = torch.autograd.grad(yr, (xr, ref.weight, ref.bias), dy.float())
    assert _rel_err(gx, gxr.cpu()) <= tol and _rel_err(gw, gwr.cpu()) <= tol and _rel_err(gb, gbr.cpu()) <= tol
    torch.testing.assert_close(ours.running_mean, ref.running_mean, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(ours.running_var, ref.running_var, rtol=1e-4, atol=1e-4)
    assert int(ours.num_batches_tracked) == 2 and set(ours.state_dict()) == set(ref.state_dict())
    ours.eval(), ref.eval()
    ye = ours(grid.jagged_like(x), grid).jdata
    yer = torch.relu(ref(xr)) if relu else ref(xr)
    assert _rel_err(ye.detach(), yer.detach().cpu()) <= tol
    (gxe,) = torch.autograd.grad(ye, x, dy)
    (gxer,) = torch.autograd.grad(yer, xr, dy.float())
    assert _rel_err(gxe, gxer.cpu()) <= tol


def test_bias_gradient_column_sums(fvdb):
    from fvdb import _norm

    gen = torch.Generator().manual_seed(2)
    for dtype, c in ((torch.float32, 32), (torch.bfloat16, 64), (torch.float16, 16), (torch.float32, 5)):
        x = torch.randn((70001, c), generator=gen).to(dtype).to(DEV)
        got = _norm.column_sums(x)
        assert got.dtype == torch.float32 and _rel_err(got, x.double().sum(0).float().cpu()) <= 1e-6


# ------------------------------------------------------------------ edge cases and full-size properties


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_empty_and_degenerate_batches(fvdb, dtype):
    # GatherScatterDefault.cu:698-699,771-777: O == 0 / P == 0 -> zeros of the right shape; empty batch items are preserved
    # (BuildGridForConv.cu:372-389); a single voxel; a target grid disjoint from the source (no pairs at all).
    cpp = fvdb._fvdb_cpp
    w = (torch.ones((32, 32, 3, 3, 3)) / 8).to(dtype).to(DEV).requires_grad_()
    # (a) batch with an empty item in the middle
    coords = [_random_batch(1, n=900, extent=8, batches=1)[0], np.zeros((0, 3), dtype=np.int64), _random_batch(2, n=700, extent=8, batches=1)[0]]
    grid = _grid(fvdb, coords)
    assert grid.grid_count == 3 and grid.num_voxels_at(1) == 0
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid, grid)
    x = torch.randn((grid.total_voxels, 32)).to(dtype).to(DEV).requires_grad_()
    y = plan.execute(grid.jagged_like(x), w)
    want = _oracle_run(plan._backend.topology, x.detach(), w.detach(), torch.ones_like(y.jdata))[0]
    assert _rel_err(y.jdata.detach(), want) <= (1e-5 if dtype == torch.float32 else 2e-2)
    assert y.joffsets.tolist() == grid.joffsets.tolist()
    strided = fvdb.ConvolutionPlan.from_grid_batch(2, 2, grid)
    assert strided.target_grid_batch.grid_count == 3 and strided.target_grid_batch.num_voxels_at(1) == 0
    # (b) every item empty
    none = _grid(fvdb, [np.zeros((0, 3), dtype=np.int64)] * 2)
    p0 = fvdb.ConvolutionPlan.from_grid_batch(3, 1, none, none)
    x0 = torch.zeros((0, 32), dtype=dtype, device=DEV, requires_grad=True)
    y0 = p0.execute(none.jagged_like(x0), w)
    assert y0.jdata.shape == (0, 32) and p0._backend.topology.total_pairs == 0
    gx0, gw0 = torch.autograd.grad(y0.jdata.sum(), (x0, w), allow_unused=True)
    assert gx0.shape == (0, 32) and float(gw0.abs().sum()) == 0.0
    # (c) one voxel: only the centre tap is live
    one = _grid(fvdb, [[(5, -3, 2)]])
    p1 = fvdb.ConvolutionPlan.from_grid_batch(3, 1, one, one)
    t1 = p1._backend.topology
    assert t1.total_pairs == 1 and t1.offsets.tolist() == [0] * 14 + [1] * 14
    x1 = torch.arange(32, dtype=torch.float32).reshape(1, 32).to(dtype).to(DEV)
    torch.testing.assert_close(p1.execute(x1, w).float(), (x1.float().sum() / 8).expand(1, 32), rtol=1e-2, atol=1e-2)
    # (d) explicit target far away from the source: RESTRICTED plan with no pair; outputs are zeros, gradients are zeros
    far = _grid(fvdb, [[(100, 100, 100), (101, 100, 100)]])
    pd = fvdb.ConvolutionPlan.from_grid_batch(3, 1, one, far)
    assert pd._backend.topology.total_pairs == 0
    xd = torch.ones((1, 32), dtype=dtype, device=DEV, requires_grad=True)
    yd = pd.execute(xd, w)
    assert yd.shape == (2, 32) and float(yd.abs().sum()) == 0.0
    gxd, gwd = cpp.gs_conv_backward(torch.ones_like(yd), xd.detach(), w.detach(), pd._backend.topology)
    assert float(gxd.abs().sum()) == 0.0 and float(gwd.abs().sum()) == 0.0


def test_full_size_properties_on_the_bench_workload(fvdb):
    # BASELINE.json configs[1] at full size (8 indoor grids x ~200 k voxels, 3^3 64->64 bf16) through size-independent
    # properties: map symmetry and per-grid isolation (bit-exact), all-ones == rulebook degree (exact in bf16), the adjoint
    # identity <conv(x), d> == <x, conv^T(d)> == <W, wgrad>, linearity, and run-to-run determinism.
    import bench

    cfg = bench.CONFIGS["c2"]
    grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(bench.make_coords(cfg, 0, torch.device(DEV))))
    n = grid.total_voxels
    assert grid.grid_count == 8 and 8 * 190_000 <= n <= 8 * 210_000
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid, grid)
    topo = plan._backend.topology
    nbr = topo._out_map()[:, :n].long()  # [27, n]
    rows = torch.arange(n, device=DEV)
    assert torch.equal(nbr[13], rows)  # centre tap = identity
    for k in (0, 5, 12):  # stride-1 same grid: nbr[k][o] = i  <=>  nbr[26-k][i] = o
        hit = nbr[k] >= 0
        assert torch.equal(nbr[26 - k][nbr[k][hit]], rows[hit])
    jidx = grid.jidx.long()
    for k in (0, 9, 26):  # the map never crosses grids (GatherScatterDefault.cu:126,186-188)
        hit = nbr[k] >= 0
        assert torch.equal(jidx[nbr[k][hit]], jidx[hit])
    degree = (nbr >= 0).sum(0)
    assert int(degree.sum()) == topo.total_pairs == int(topo.offsets[-1])
    ones_x = torch.ones((n, 64), dtype=torch.bfloat16, device=DEV)
    ones_w = torch.full((64, 64, 3, 3, 3), 1.0 / 64, dtype=torch.bfloat16, device=DEV)
    y = plan.execute(ones_x, ones_w) if grid.grid_count == 1 else plan.execute(grid.jagged_like(ones_x), ones_w).jdata
    assert torch.equal(y.float(), degree.float()[:, None].expand(n, 64))  # sums <= 27 are exact in bf16
    gen = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn((n, 64), generator=gen, device=DEV).bfloat16()
    d = torch.randn((n, 64), generator=gen, device=DEV).bfloat16()
    w = (torch.randn((64, 64, 3, 3, 3), generator=gen, device=DEV) / 41.6).bfloat16()
    cpp = fvdb._fvdb_cpp
    yx = cpp.gs_conv(x, w, topo)
    gx, gw = cpp.gs_conv_backward(d, x, w, topo)
    lhs = float((yx.double() * d.double()).sum())
    assert abs(lhs - float((x.double() * gx.double()).sum())) <= 2e-3 * abs(lhs)
    assert abs(lhs - float((w.double() * gw.double()).sum())) <= 2e-3 * abs(lhs)
    y2 = cpp.gs_conv(x, (2 * w.float()).bfloat16(), topo)  # scaling by a power of two is exact
    assert torch.equal(y2.float(), 2 * yx.float())
    again = cpp.gs_conv_backward(d, x, w, topo)
    assert torch.equal(cpp.gs_conv(x, w, topo), yx) and torch.equal(again[0], gx) and torch.equal(again[1], gw)  # no atomics anywhere


# ------------------------------------------------------------------ pooling / refinement


@pytest.mark.parametrize("dtype,channels", [(torch.float32, 12), (torch.bfloat16, 32)])
@pytest.mark.parametrize("factor", [2, (2, 3, 1)])
def test_pool_refine_match_oracle(fvdb, dtype, channels, factor):
    from oracle import pool_oracle as po

    f = oracle.normalize_3d(factor)
    coords = _random_batch(70, n=2500, extent=9, batches=2)
    fine = _grid(fvdb, coords, voxel_sizes=(0.5, 1.0, 2.0), origins=(3.0, -2.0, 7.0))
    rows, bidx = _rows(fine)
    coarse = fine.coarsened_grid(factor)
    crow, cb = _rows(coarse)
    for b in range(2):  # coordinate sets, metadata
        want = po.coarsened_ijk(rows[bidx == b], f)
        assert sorted(map(tuple, crow[cb == b].tolist())) == sorted(map(tuple, want.tolist()))
    ws, wo = po.coarse_metadata((0.5, 1.0, 2.0), (3.0, -2.0, 7.0), f)
    torch.testing.assert_close(coarse.voxel_sizes.cpu().double(), torch.tensor([ws, ws])), torch.testing.assert_close(coarse.origins.cpu().double(), torch.tensor([wo, wo]))
    gen = torch.Generator().manual_seed(8)
    x = torch.randn((fine.total_voxels, channels), generator=gen).to(dtype)
    dy = torch.randn((coarse.total_voxels, channels), generator=gen).to(dtype)
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    for mode, fn in (("max", fine.max_pool), ("avg", fine.avg_pool)):
        xd = x.to(DEV).requires_grad_()
        y, cg = fn(factor, fine.jagged_like(xd))
        assert cg.total_voxels == coarse.total_voxels and y.jdata.dtype == dtype
        want_y, children = po.pool(rows, bidx, x.double().numpy(), crow, cb, f, (0, 0, 0), mode)
        assert _rel_err(y.jdata.detach(), torch.from_numpy(want_y).float()) <= tol
        (gx,) = torch.autograd.grad(y.jdata, xd, dy.to(DEV))
        want_gx = po.pool_backward(dy.double().numpy(), x.double().numpy(), children, len(rows), mode)
        assert _rel_err(gx, torch.from_numpy(want_gx).float()) <= tol
        y2, _ = fn(factor, fine.jagged_like(xd), coarse_grid=coarse)  # explicit coarse grid: same rows
        assert torch.equal(y2.jdata, y.jdata)
    # refine back onto the original fine grid, and onto the generated refined grid
    z = torch.randn((coarse.total_voxels, channels), generator=gen).to(dtype)
    zd = z.to(DEV).requires_grad_()
    up, fg = coarse.refine(factor, coarse.jagged_like(zd), fine_grid=fine)
    want_up, parent = po.refine(crow, cb, z.double().numpy(), rows, bidx, f)
    assert fg.is_same(fine) and (parent >= 0).all() and _rel_err(up.jdata.detach(), torch.from_numpy(want_up).float()) <= tol
    dup = torch.randn((fine.total_voxels, channels), generator=gen).to(dtype)
    (gz,) = torch.autograd.grad(up.jdata, zd, dup.to(DEV))
    assert _rel_err(gz, torch.from_numpy(po.refine_backward(dup.double().numpy(), parent, len(crow))).float()) <= tol
    refined = coarse.refined_grid(factor)
    rrow, rb = _rows(refined)
    for b in range(2):
        want = po.refined_ijk(crow[cb == b], f)
        assert sorted(map(tuple, rrow[rb == b].tolist())) == sorted(map(tuple, want.tolist()))
    fs, fo = po.fine_metadata(ws, wo, f)
    torch.testing.assert_close(refined.voxel_sizes.cpu().double(), torch.tensor([fs, fs])), torch.testing.assert_close(refined.origins.cpu().double(), torch.tensor([fo, fo]))
    masked = coarse.refined_grid(factor, mask=coarse.jagged_like(torch.arange(coarse.total_voxels, device=DEV) % 2 == 0))
    assert masked.total_voxels == ((coarse.total_voxels + 1) // 2) * f[0] * f[1] * f[2]
    with pytest.raises(ValueError, match="must not overlap"):
        fine.max_pool(3, fine.jagged_like(x.to(DEV)), stride=2)


def test_nn_pooling_modules_and_unet_style_block(fvdb):
    # MaxPool -> 1x1x1 conv -> BatchNorm(+ReLU) -> UpsamplingNearest, the down / up pattern of the reference's SimpleUNet
    # (fvdb/nn/simple_unet.py:194-294), trained for one step.
    fine = _grid(fvdb, _random_batch(90, n=4000, extent=12, batches=2))
    coarse = fine.coarsened_grid(2)
    pool, up = fvdb.nn.MaxPool(2), fvdb.nn.UpsamplingNearest(2)
    fan_out = fvdb.nn.SparseConv3d(16, 32, kernel_size=1, bias=False).to(DEV)
    norm = fvdb.nn.BatchNorm(32, activation="relu").to(DEV)
    plan = fvdb.ConvolutionPlan.from_grid_batch(1, 1, coarse, coarse)
    x = fine.jagged_like(torch.randn((fine.total_voxels, 16), device=DEV, requires_grad=True))
    pooled, cg = pool(x, fine, coarse)
    assert cg.is_same(coarse) and pooled.jdata.shape == (coarse.total_voxels, 16)
    h = norm(fan_out(pooled, plan), coarse)
    out, fg = up(h, coarse, fine_grid=fine)
    assert fg.is_same(fine) and out.jdata.shape == (fine.total_voxels, 32)
    out.jdata.square().mean().backward()
    assert torch.isfinite(x.jdata.grad).all() and float(fan_out.weight.grad.abs().sum()) > 0 and float(norm.weight.grad.abs().sum()) > 0
    avg, _ = fvdb.nn.AvgPool(2)(x, fine)
    assert avg.jdata.shape == (coarse.total_voxels, 16)


@pytest.mark.parametrize("dtype,widths", [(torch.float32, (3, 4, 5)), (torch.bfloat16, (8, 16, 8))])
def test_simple_unet_forward_backward(fvdb, dtype, widths):
    # The reference's own SimpleUNet checks (tests/unit/test_simple_unet.py:33-160): dense 16^3 block plus a sparse second
    # grid, finite outputs of the right shape, finite gradients for every parameter, reference-compatible state_dict keys.
    cin, base, cout = widths
    dense = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij"), -1).reshape(-1, 3)
    grid = _grid(fvdb, [dense, _random_batch(33, n=1500, extent=7, batches=1)[0]])
    torch.manual_seed(42)
    model = fvdb.nn.SimpleUNet(cin, base, cout, channel_growth_rate=2, kernel_size=3, downup_layer_count=2, block_layer_count=1).to(DEV).to(dtype)
    keys = set(model.state_dict())
    for expected in ("pad.conv.weight", "pad.batch_norm.running_var", "downup.conv_in.blocks.0.conv.weight", "downup.down.channel_fan_out.weight",
                     "downup.inner.down.batch_norm.weight", "downup.inner.inner.block.blocks.0.batch_norm.bias", "downup.inner.up.channel_fan_in.weight",
                     "downup.up.batch_norm.num_batches_tracked", "downup.conv_out.blocks.0.conv.weight", "unpad.deconv.weight"):
        assert expected in keys
    assert model.downup.down.channel_fan_out.weight.shape == (2 * base, base) and model.unpad.deconv.weight.shape == (cout, base, 3, 3, 3)
    x = grid.jagged_like(torch.randn((grid.total_voxels, cin), device=DEV).to(dtype))
    out = model(x, grid)
    assert out.jdata.shape == (grid.total_voxels, cout) and out.jdata.dtype == dtype and torch.isfinite(out.jdata).all()
    out.jdata.float().sum().backward()
    for name, param in model.named_parameters():
        assert param.grad is not None and torch.isfinite(param.grad).all(), name
    single = fvdb.nn.SimpleUNet(cin, base, cout, 2, downup_layer_count=1, block_layer_count=2).to(DEV).to(dtype)
    with torch.no_grad():
        assert torch.isfinite(single(x, grid).jdata).all()


@pytest.mark.parametrize("ks", [2, 3, 5, (3, 5, 7), 8, (1, 4, 2)])
def test_stride1_leaf_morphology_equals_candidate_path(fvdb, ks):
    # conv_grid / conv_transpose_grid at stride 1 by leaf-mask dilation (csrc/grid_morph.cu) == the sort-unique candidate path,
    # voxel for voxel and row for row; the rebuilt leaves answer lookups consistently (ijk_to_index o ijk == identity).
    cpp = fvdb._fvdb_cpp
    from fvdb.utils.synthetic import sphere_shell

    coords = [sphere_shell(target=5000, domain=64, seed=5, device="cpu").numpy() - 30, _random_batch(61, n=2500, extent=25, batches=1)[0],
              np.array([[4095, 4095, 4095], [4096, 4096, 4096], [-4097, 7, 8], [-1, -1, -1]])]  # straddles root tiles
    grid = _grid(fvdb, coords)
    for fn in (grid.conv_grid, grid.conv_transpose_grid):
        try:
            cpp.use_leaf_morphology = False
            want = fn(ks, 1)
        finally:
            cpp.use_leaf_morphology = True
        got = fn(ks, 1)
        assert torch.equal(got.ijk.jdata, want.ijk.jdata) and torch.equal(got.jidx, want.jidx) and torch.equal(got.joffsets, want.joffsets)
        idx = got.ijk_to_index(got.ijk, cumulative=True).jdata
        assert torch.equal(idx, torch.arange(got.total_voxels, device=DEV))
        probe = got.ijk.jdata + torch.tensor([[0, 0, 9]], device=DEV, dtype=torch.int32)  # mostly outside: agree with the reference build
        assert torch.equal(got.ijk_to_index(got.jagged_like(probe)).jdata, want.ijk_to_index(want.jagged_like(probe)).jdata)
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid)  # target generated by the fast path
    assert plan._backend.topology.total_pairs == 27 * grid.total_voxels  # complete support: every tap of every voxel lands
