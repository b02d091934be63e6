"""Value parity on the BASELINE.json configurations themselves, at size and on their own generators.

Every case builds the kernel map with the CPU oracle FROM ``grid.ijk`` (never from the GPU's map), checks the GPU map
against it bit for bit, then compares ``y``, ``grad_x`` and ``grad_w`` of the CUDA path (through the C ABI) with the
oracle's per-tap gather -> mm -> scatter-add restatement (GatherScatterDefault.cu:706-721,786-813):

* C1  100 k voxels, 3^3 32->32 fp32: elementwise, the reference's own fp32 bar (rtol 1e-5 / atol 1e-6 for forward and
  input gradient, 5e-4 / 5e-4 for the kernel gradient: fvdb/utils/tests/convolution_utils.py:87-137) against the
  fp64-accumulated oracle;
* C2  8 x ~205 k voxels, 3^3 64->64 bf16 (full size): 2e-2 against the fp32 oracle (north_star), norm-wise and elementwise;
* C4  one ~1 M-voxel LiDAR grid, 3^3 128->128 bf16;
* C5  one ~5 M-voxel 20 %-occupancy grid, 5^3 16->16 bf16;
* C3  the UNet block stack layer by layer (3^3 same-topology, 2^3 stride-2 down on the generated target,
  from_plan_transposed up; 32..256 channels) on indoor grids;
* by-grid partition: the batch split over two devices reproduces the one-device ``y`` / ``grad_x`` bit for bit and the
  summed ``grad_w`` to fp32 round-off (skipped with fewer than 2 GPUs).

Mirrors /root/reference/tests/unit/test_conv_semantics_integration.py:170-242 (production against an independent oracle on
the same inputs) at the sizes the benchmark is quoted on.
"""

import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"
REPO = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def fvdb():
    import fvdb as module

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return module


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", REPO / "bench.py")
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def _rows(grid):
    return grid.ijk.jdata.cpu().numpy().astype(np.int64), grid.jidx.cpu().numpy().astype(np.int64)


def _oracle_topology(source, target, ks, st, transposed=False):
    s_ijk, s_b = _rows(source)
    t_ijk, t_b = _rows(target)
    return oracle.build_topology(s_ijk, s_b, t_ijk, t_b, ks, st, transposed=transposed)


def _assert_map_equals_oracle(topo, ref):
    """Both sides order a tap segment by output row and an output row has at most one pair per tap: plain equality."""
    assert topo.offsets.tolist() == ref.offsets.tolist()
    assert np.array_equal(topo.gather_indices.cpu().numpy(), ref.gather_indices)
    assert np.array_equal(topo.scatter_indices.cpu().numpy(), ref.scatter_indices)


def _inputs(n_in, n_out, cin, cout, k, dtype, seed):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn((n_in, cin), generator=gen).to(dtype)
    bound = 1.0 / (cin * k[0] * k[1] * k[2]) ** 0.5  # fvdb/nn/modules.py:304-309
    w = ((torch.rand((cout, cin, *k), generator=gen) * 2 - 1) * bound).to(dtype)
    dy = torch.randn((n_out, cout), generator=gen).to(dtype)
    return x, w, dy


def _check_half(name, got, want, tol=2e-2):
    got = got.float().cpu()
    rel = float((got - want).norm() / want.norm().clamp_min(1e-30))
    assert rel <= tol, f"{name}: relative error {rel:.3e} > {tol}"
    # elementwise: bf16 rounding of the result (2^-9 relative) plus products rounded nowhere else -> well inside 2e-2 of the
    # value plus 2e-2 of the typical magnitude
    rms = float(want.pow(2).mean().sqrt())
    bad = (got - want).abs() > tol * want.abs() + tol * rms
    assert not bool(bad.any()), f"{name}: {int(bad.sum())} of {bad.numel()} elements beyond {tol} (max abs err {float((got - want).abs().max()):.3e}, rms {rms:.3e})"


def _run_gpu(fvdb, topo, x, w, dy):
    cpp = fvdb._fvdb_cpp
    xd, wd, dyd = x.to(DEV), w.to(DEV), dy.to(DEV)
    fwd = cpp.gs_conv_transpose if topo.is_transposed else cpp.gs_conv
    bwd = cpp.gs_conv_transpose_backward if topo.is_transposed else cpp.gs_conv_backward
    y = fwd(xd, wd, topo)
    gx, gw = bwd(dyd, xd, wd, topo)
    return y, gx, gw


def _compare_layer(fvdb, plan, source, target, ks, st, cin, cout, dtype, seed, transposed=False, ref=None):
    topo = plan._backend.topology
    ref = ref if ref is not None else _oracle_topology(source, target, ks, st, transposed)
    _assert_map_equals_oracle(topo, ref)
    k = oracle.normalize_3d(ks)
    x, w, dy = _inputs(source.total_voxels, target.total_voxels, cin, cout, k, dtype, seed)
    y, gx, gw = _run_gpu(fvdb, topo, x, w, dy)
    acc = torch.float64 if dtype == torch.float32 else torch.float32
    want_y = oracle.gs_conv(x.to(acc), w.to(acc), ref, accumulate_dtype=acc)
    want_gx, want_gw = oracle.gs_conv_backward(dy.to(acc), x.to(acc), w.to(acc), ref, accumulate_dtype=acc)
    if dtype == torch.float32:
        torch.testing.assert_close(y.cpu().double(), want_y, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(gx.cpu().double(), want_gx, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(gw.cpu().double(), want_gw, rtol=5e-4, atol=5e-4)
    else:
        _check_half("y", y, want_y)
        _check_half("grad_x", gx, want_gx)
        _check_half("grad_w", gw, want_gw)
    return ref


def test_c1_fp32_100k_elementwise_reference_tolerances(fvdb, bench):
    cfg = bench.CONFIGS["c1"]
    grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(bench.make_coords(cfg, 0, torch.device(DEV))))
    assert 98_000 <= grid.total_voxels <= 102_000
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid, grid)
    _compare_layer(fvdb, plan, grid, grid, 3, 1, 32, 32, torch.float32, seed=1)


def test_c2_full_size_bf16_all_of_y_grad_x_grad_w(fvdb, bench):
    cfg = bench.CONFIGS["c2"]
    grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(bench.make_coords(cfg, 0, torch.device(DEV))))
    assert grid.grid_count == 8 and 8 * 190_000 <= grid.total_voxels <= 8 * 210_000
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid, grid)
    ref = _compare_layer(fvdb, plan, grid, grid, 3, 1, 64, 64, torch.bfloat16, seed=2)
    # the same full-size batch through every pipeline shape of the forward kernel (bench knob): all must agree with the oracle
    cpp = fvdb._fvdb_cpp
    x, w, _ = _inputs(grid.total_voxels, grid.total_voxels, 64, 64, (3, 3, 3), torch.bfloat16, seed=3)
    want = oracle.gs_conv(x.float(), w.float(), ref, accumulate_dtype=torch.float32)
    try:
        for variant in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10):
            cpp.set_kernel_variant(variant)
            _check_half(f"y (variant {variant})", cpp.gs_conv(x.to(DEV), w.to(DEV), plan._backend.topology), want)
    finally:
        cpp.set_kernel_variant(0)


def test_c4_one_lidar_grid_128_channels(fvdb, bench):
    cfg = dict(bench.CONFIGS["c4"], grids=1, partition=None)
    grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(bench.make_coords(cfg, 0, torch.device(DEV))))
    assert 950_000 <= grid.total_voxels <= 1_060_000
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, grid, grid)
    _compare_layer(fvdb, plan, grid, grid, 3, 1, 128, 128, torch.bfloat16, seed=4)


def test_c5_one_5m_grid_5x5x5_16_channels(fvdb, bench):
    cfg = dict(bench.CONFIGS["c5"], grids=1)
    grid = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(bench.make_coords(cfg, 0, torch.device(DEV))))
    assert 4_900_000 <= grid.total_voxels <= 5_050_000
    plan = fvdb.ConvolutionPlan.from_grid_batch(5, 1, grid, grid)
    _compare_layer(fvdb, plan, grid, grid, 5, 1, 16, 16, torch.bfloat16, seed=5)


def test_c3_block_stack_layer_by_layer(fvdb, bench):
    # four of the sixteen indoor grids keep the CPU oracle within seconds; the layers are the C3 stack's
    cfg = dict(bench.CONFIGS["c3"], grids=4)
    g0 = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(bench.make_coords(cfg, 0, torch.device(DEV))))
    widths = [32, 64, 128, 256]
    Plan = fvdb.ConvolutionPlan
    grids, seed = [g0], 10
    for level in range(4):
        g = grids[level]
        c = widths[level]
        _compare_layer(fvdb, Plan.from_grid_batch(3, 1, g, g), g, g, 3, 1, c, c, torch.bfloat16, seed=seed)
        seed += 1
        if level == 3:
            break
        down = Plan.from_grid_batch(2, 2, g)  # target = conv_grid(2, 2)
        coarse = down.target_grid_batch
        c_ijk, c_b = _rows(coarse)
        want_ijk, want_b = oracle.conv_grid(*_rows(g), 2, 2)
        order = oracle.index_grid_row_order(want_b, want_ijk)
        assert np.array_equal(c_ijk, want_ijk[order]) and np.array_equal(c_b, want_b[order])  # generated target, incl. row order
        ref_down = _compare_layer(fvdb, down, g, coarse, 2, 2, c, widths[level + 1], torch.bfloat16, seed=seed)
        seed += 1
        up = Plan.from_plan_transposed(down)  # exact adjoint topology: coarse -> fine (convolution_plan.py:837)
        ref_up = oracle.reverse_topology(ref_down)
        assert up._backend.topology.is_transposed
        _compare_layer(fvdb, up, coarse, g, 2, 2, widths[level + 1], c, torch.bfloat16, seed=seed, transposed=True, ref=ref_up)
        seed += 1
        grids.append(coarse)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_by_grid_partition_over_two_devices_reproduces_one_device(fvdb, bench):
    from fvdb.distributed import partition_grids_lpt

    cfg = bench.CONFIGS["c2"]
    coords = bench.make_coords(cfg, 0, torch.device("cuda:0"))
    whole = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(coords))
    plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, whole, whole)
    n = whole.total_voxels
    x, w, dy = _inputs(n, n, 64, 64, (3, 3, 3), torch.bfloat16, seed=6)
    cpp = fvdb._fvdb_cpp
    y1 = cpp.gs_conv(x.to("cuda:0"), w.to("cuda:0"), plan._backend.topology)
    gx1, gw1 = cpp.gs_conv_backward(dy.to("cuda:0"), x.to("cuda:0"), w.to("cuda:0"), plan._backend.topology)
    offsets = whole.joffsets.cpu().tolist() if hasattr(whole, "joffsets") else np.concatenate([[0], np.cumsum([len(c) for c in coords])]).tolist()
    shares = partition_grids_lpt([len(c) for c in coords], 2)
    assert sorted(shares[0] + shares[1]) == list(range(8))
    gw_sum = torch.zeros(w.shape, dtype=torch.float64)
    for rank, mine in enumerate(shares):
        dev = torch.device("cuda", rank)
        part = fvdb.GridBatch.from_ijk(fvdb.JaggedTensor([coords[g].to(dev) for g in mine]))
        rows = torch.cat([torch.arange(offsets[g], offsets[g + 1]) for g in mine])
        sub_plan = fvdb.ConvolutionPlan.from_grid_batch(3, 1, part, part)
        topo = sub_plan._backend.topology
        with torch.cuda.device(dev):
            y = cpp.gs_conv(x[rows].to(dev), w.to(dev), topo)
            gx, gw = cpp.gs_conv_backward(dy[rows].to(dev), x[rows].to(dev), w.to(dev), topo)
        assert torch.equal(y.cpu(), y1.cpu()[rows]) and torch.equal(gx.cpu(), gx1.cpu()[rows])  # whole grids per rank: no row sees another rank
        gw_sum += gw.cpu().double()
    # grad_w: each rank rounds its partial to bf16 before the SUM all-reduce; the one-device run rounds once
    torch.testing.assert_close(gw_sum.float(), gw1.cpu().float(), rtol=2e-2, atol=2e-2 * float(gw1.float().abs().mean()))
