"""CPU-side tests (no GPU): the C-ABI library loads and exports every declared symbol, host-only geometry entry
points, JaggedTensor, plan policy helpers, grid partitioning and the world_size-2 gloo gradient all-reduce."""

import os
import re
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO = Path(__file__).resolve().parent.parent


def test_abi_exports_every_declared_symbol():
    import ctypes

    from fvdb import _lib

    header = (REPO / "include" / "fvdbconv.h").read_text()
    declared = set(re.findall(r"FVC_API\s+[\w\s\*]+?\b(fvc_\w+)\s*\(", header))
    assert len(declared) >= 25
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(str(REPO / "fvdb-core_b200" / "fvdb" / "libfvdbconv.so"))
    for name in declared:
        assert hasattr(lib, name), f"libfvdbconv.so does not export {name}"
    assert _lib.lib.fvc_abi_version() == _lib.ABI_VERSION == 3


def test_geometry_entry_points_match_oracle_and_reference_kats():
    import fvdb
    import oracle

    g = fvdb._fvdb_cpp.ConvolutionGeometry([4, 3, 6], [2, 3, 4])
    assert (g.kernel_volume, g.padding_before, g.padding_after) == (72, [1, 1, 2], [2, 1, 3])
    assert g.tap_coord(23) == [1, 0, 5] and g.fine_from_coarse([3, -2, 1], [0, 0, 0]) == [5, -7, 2]
    assert g.semantics_version == 1 and g.phase_policy == "torch_same_phase" and g.dilation == [1, 1, 1] and g.registration_offset == [0, 0, 0]
    g4 = fvdb._fvdb_cpp.ConvolutionGeometry([4, 4, 4], [4, 4, 4])
    assert g4.coarse_from_fine([-2, -2, -2], [3, 3, 3]) == [-1, -1, -1] and g4.coarse_from_fine([-2, -2, -2], [2, 2, 2]) is None
    rng = np.random.default_rng(0)
    for _ in range(200):
        ks, st = rng.integers(1, 7, 3).tolist(), rng.integers(1, 6, 3).tolist()
        geo, ref = fvdb._fvdb_cpp.ConvolutionGeometry(ks, st), oracle.Geometry(ks, st)
        tap = ref.tap_coord(int(rng.integers(0, ref.kernel_volume)))
        c = rng.integers(-50, 50, 3).tolist()
        assert geo.fine_from_coarse(c, tap) == ref.fine_from_coarse(c, tap).tolist()
        coarse, ok = ref.coarse_from_fine(c, tap)
        assert geo.coarse_from_fine(c, tap) == (coarse.tolist() if bool(ok) else None)
    with pytest.raises(ValueError, match="kernel_size must be strictly positive"):
        fvdb._fvdb_cpp.ConvolutionGeometry([0, 1, 1], [1, 1, 1])
    with pytest.raises(ValueError, match="stride must be strictly positive"):
        fvdb._fvdb_cpp.ConvolutionGeometry([1, 1, 1], [1, -2, 1])
    with pytest.raises(IndexError):
        g.tap_coord(72)


def test_no_cpu_fallback():
    import fvdb

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fvdb.GridBatch.from_ijk(fvdb.JaggedTensor(torch.zeros((4, 3), dtype=torch.int32)))


def test_jagged_tensor_subset():
    from fvdb import JaggedTensor

    a, b = torch.arange(6.0).reshape(3, 2), torch.arange(4.0).reshape(2, 2) + 10
    jt = JaggedTensor([a, b])
    assert jt.num_tensors == 2 and jt.joffsets.tolist() == [0, 3, 5] and jt.jidx.tolist() == [0, 0, 0, 1, 1] and jt.jidx.dtype == torch.int32
    assert torch.equal(jt[1].jdata, b) and [t.shape[0] for t in jt.unbind()] == [3, 2] and jt.lshape == [3, 2]
    like = jt.jagged_like(torch.zeros(5, 7))
    assert like.rshape == (5, 7) and like.joffsets.tolist() == [0, 3, 5]
    with pytest.raises(ValueError):
        jt.jagged_like(torch.zeros(4, 7))
    jt.jdata = jt.jdata + 1
    assert float(jt.jdata[0, 0]) == 1.0
    assert torch.equal((jt * 2).jdata, jt.jdata * 2)
    single = JaggedTensor(a)
    assert single.num_tensors == 1 and single.joffsets.tolist() == [0, 3]
    off = JaggedTensor.from_data_and_offsets(torch.zeros(5, 1), torch.tensor([0, 0, 5]))
    assert off.num_tensors == 2 and off.jidx.tolist() == [1] * 5
    with pytest.raises(ValueError):
        JaggedTensor([a, torch.zeros(2, 3)])


def test_plan_policy_helpers_and_types():
    import fvdb
    from fvdb import convolution_plan as cp
    from fvdb.types import ValueConstraint, to_Vec3i

    assert to_Vec3i(3).tolist() == [3, 3, 3] and to_Vec3i([1, 2, 3]).tolist() == [1, 2, 3]
    with pytest.raises(ValueError):
        to_Vec3i(0, value_constraint=ValueConstraint.POSITIVE)
    with pytest.raises(TypeError):
        to_Vec3i(1.5)
    P = fvdb.ConvolutionTopologyPolicy
    assert cp._resolve_topology_policy(None, None) is P.COMPLETE and cp._resolve_topology_policy(object(), None) is P.RESTRICTED
    with pytest.raises(ValueError, match="COMPLETE.*target_grid=None"):
        cp._resolve_topology_policy(object(), P.COMPLETE)
    with pytest.raises(TypeError):
        cp._resolve_topology_policy(None, "complete")
    cp._WARNED_INCOMPLETE_COVERAGE_GEOMETRIES.clear()
    geo = fvdb._fvdb_cpp.ConvolutionGeometry([1, 1, 1], [2, 2, 2])
    with pytest.warns(fvdb.ConvolutionCoverageWarning, match="uncovered stride residues"):
        cp._warn_if_incomplete_residue_coverage(geo, False)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("error")
        cp._warn_if_incomplete_residue_coverage(geo, False)  # once per geometry
        cp._warn_if_incomplete_residue_coverage(fvdb._fvdb_cpp.ConvolutionGeometry([3, 3, 3], [2, 2, 2]), False)  # full coverage
    with pytest.raises(ValueError, match="uniform kernel sizes 3, 5, 7"):
        cp._validate_pred_gather_igemm_admission(torch.tensor([3, 3, 5]), torch.tensor([1, 1, 1]), (), transposed=False)
    with pytest.raises(ValueError, match="uniform strides 1, 2"):
        cp._validate_pred_gather_igemm_admission(torch.tensor([3, 3, 3]), torch.tensor([3, 3, 3]), (), transposed=False)
    assert cp._matmul_weight_matrix(torch.zeros(4, 2, 1, 1, 1)).shape == (4, 2)
    m = fvdb.nn.SparseConv3d(4, 8, 3)
    assert m.weight.shape == (8, 4, 3, 3, 3) and m.weight.stride() == (1, 8, 32, 96, 288) and m.bias.shape == (8,)
    assert fvdb.nn.SparseConv3d(4, 8, 1).weight.shape == (8, 4)
    bound = 1 / (4 * 27) ** 0.5
    assert float(m.weight.abs().max()) <= bound and set(m.state_dict()) == {"weight", "bias"}


def test_partition_grids_lpt():
    from fvdb.distributed import partition_grids_lpt

    counts = [1000, 10, 900, 20, 500, 480, 30, 5]
    for world in (1, 2, 3, 8):
        bins = partition_grids_lpt(counts, world)
        assert sorted(g for b in bins for g in b) == list(range(8)) and len(bins) == world
    two = partition_grids_lpt(counts, 2)
    loads = [sum(counts[g] for g in b) for b in two]
    assert abs(loads[0] - loads[1]) <= 100
    assert partition_grids_lpt([5, 5], 4) == [[0], [1], [], []]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _allreduce_worker(rank: int, world: int, port: int, out_dir: str):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, str(REPO / "fvdb-core_b200"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fvdb.distributed import allreduce_gradients, partition_grids_lpt

    torch.manual_seed(0)
    weights = [torch.nn.Parameter(torch.zeros(8, 4, 3, 3, 3)), torch.nn.Parameter(torch.zeros(8)), torch.nn.Parameter(torch.zeros(3, dtype=torch.float64))]
    mine = partition_grids_lpt([7, 3, 5, 1], world)[rank]  # each rank owns whole grids
    for p in weights:  # a rank's "wgrad" = sum over its own grids of a per-grid contribution
        p.grad = sum((torch.full_like(p, float(g + 1)) for g in mine), torch.zeros_like(p))
    calls = allreduce_gradients(weights, bucket_bytes=1 << 12)
    assert calls >= 2
    for p in weights:
        assert torch.equal(p.grad, torch.full_like(p, float(1 + 2 + 3 + 4)))  # identical to the single-process sum
    dist.barrier()
    dist.destroy_process_group()
    Path(out_dir, f"ok{rank}").write_text("ok")


def test_wgrad_allreduce_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_allreduce_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def _syncbn_worker(rank, world, port, out_dir):
    import os

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fvdb

    torch.manual_seed(3)
    rows = torch.randn(300, 12, dtype=torch.float64) * 3 + 50  # |mean| >> std
    # rank 1 owns ZERO rows (fewer grids than ranks under the by-grid partition): it must still join every collective
    mine = rows if rank == 0 else rows[:0]
    bn = fvdb.nn.SyncBatchNorm(12, activation="relu").double()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5), bn.bias.uniform_(-0.5, 0.5)
    x = mine.clone().requires_grad_()
    jt = fvdb.JaggedTensor.from_data_and_indices(x, torch.zeros(len(mine), dtype=torch.int32), 1)
    y = bn(jt).jdata
    loss = y.square().sum()
    loss.backward()
    ref = torch.nn.BatchNorm1d(12).double()
    ref.load_state_dict({k: v.clone() for k, v in bn.state_dict().items() if k in ("weight", "bias")}, strict=False)
    xr = rows.clone().requires_grad_()
    yr = torch.relu(ref(xr))
    yr.square().sum().backward()
    if rank == 0:
        torch.testing.assert_close(y, yr, rtol=1e-9, atol=1e-9)
        torch.testing.assert_close(x.grad, xr.grad, rtol=1e-8, atol=1e-8)
        torch.testing.assert_close(bn.running_var, ref.running_var, rtol=1e-9, atol=1e-9)
    else:
        assert y.shape == (0, 12) and x.grad.shape == (0, 12)
    dist.barrier()
    dist.destroy_process_group()
    Path(out_dir, f"bn{rank}").write_text("ok")


def test_sync_batch_norm_with_an_empty_rank_world_size_2_gloo(tmp_path):
    port = _free_port()
    mp.spawn(_syncbn_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "bn0").exists() and (tmp_path / "bn1").exists()


def test_group_norm_matches_per_grid_torch_group_norm():
    # fvdb.nn.GroupNorm == torch GroupNorm applied to each grid's [1, C, N_b] slab (reference modules.py:452-480); pure torch
    # composition, so it runs on the CPU with a stand-in for the grid.
    import fvdb

    class _Grid:
        grid_count = 3
        jidx = torch.tensor([0] * 50 + [1] * 7 + [2] * 120, dtype=torch.int32)

        def jagged_like(self, data):
            return fvdb.JaggedTensor.from_data_and_indices(data, self.jidx, 3)

    torch.manual_seed(0)
    x = torch.randn(177, 12, dtype=torch.float64).float().requires_grad_()
    gn = fvdb.nn.GroupNorm(4, 12)
    with torch.no_grad():
        gn.weight.uniform_(0.5, 1.5), gn.bias.uniform_(-0.5, 0.5)
    grid = _Grid()
    out = gn(grid.jagged_like(x), grid).jdata
    want = torch.cat([torch.nn.functional.group_norm(x[grid.jidx == b].T[None], 4, gn.weight, gn.bias, gn.eps)[0].T for b in range(3)])
    torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)
    (g,) = torch.autograd.grad(out.square().sum(), x)
    (gw,) = torch.autograd.grad(want.square().sum(), x)
    torch.testing.assert_close(g, gw, rtol=1e-4, atol=1e-5)


def test_bench_reference_arm_prints_the_contract_line():
    # `bench.py --impl reference` needs no GPU (it times the oracle port on the host cores); guard its JSON contract here.
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--config", "c1", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, check=True).stdout.strip().splitlines()[-1]
    line = json.loads(out)
    assert line["impl"] == "reference" and line["metric"] == "sparse-conv voxels/sec fwd+bwd" and line["unit"] == "voxels/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "sample" in line["config"]


def test_bench_partition_by_grid_is_a_disjoint_cover():
    # C4-style strong scaling: every rank derives the same LPT partition of the one batch and keeps its own grids.
    sys_path_bench = str(REPO)
    import sys

    if sys_path_bench not in sys.path:
        sys.path.insert(0, sys_path_bench)
    import bench

    cfg = dict(gen="sphere_shell", grids=5, voxels=1500, kernel=3, cin=16, cout=16, dtype="bf16", partition="by_grid", desc="test")
    whole = bench.make_coords({**cfg, "partition": None}, 0, "cpu", 1)
    parts = [bench.make_coords(cfg, rank, "cpu", 2) for rank in range(2)]
    assert sum(len(p) for p in parts) == 5 and all(len(p) >= 2 for p in parts)
    seen = sorted(int(c.shape[0]) for p in parts for c in p)
    assert seen == sorted(int(c.shape[0]) for c in whole)
    loads = [sum(int(c.shape[0]) for c in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= max(int(c.shape[0]) for c in whole)  # LPT: imbalance bounded by one grid


def test_bench_rooflines_credit_the_fused_backward_with_the_passes_it_replaces():
    # the step roofline is the sum of the per-pass gathered bytes, whether the backward runs as two
    # kernels or as the one fused kernel of the narrow layers; the compulsory floor never exceeds the gathered-bytes figure.
    if str(REPO) not in sys.path:
        sys.path.insert(0, str(REPO))
    import bench

    P, n, cin, cout, k3, s = 1_000_000_000, 40_000_000, 16, 16, 125, 2
    peaks = {"hbm_gbs": 6546.9, "tflops": 1346.6, "source": "test"}
    ab = bench.algorithmic_bytes(P, n, n, cin, cout, k3, s)
    assert ab["fwd"] == P * cin * s + n * cout * s + 4 * P + k3 * cin * cout * s
    assert ab["wgrad"] == P * (cin + cout) * s + 8 * P + 4 * k3 * cin * cout
    assert ab["bwd_fused"] == ab["dgrad"] + ab["wgrad"]
    assert bench.fused_formulation_bytes(P, n, n, cin, cout, k3, s) < ab["bwd_fused"]
    separate, roof_sep, comp_sep = bench.kernel_rooflines({"fwd": 13.5, "dgrad": 13.5, "wgrad": 16.8}, P, n, n, cin, cout, k3, s, peaks)
    fused, roof_fused, comp_fused = bench.kernel_rooflines({"fwd": 13.5, "dgrad": 13.5, "wgrad": 16.8, "bwd_fused": 24.0}, P, n, n, cin, cout, k3, s, peaks)
    assert abs(roof_sep - roof_fused) < 1e-9  # same algorithm, same step roofline
    assert comp_fused <= comp_sep and comp_sep < roof_sep
    assert fused["bwd_fused"]["flops"] == 2 * fused["dgrad"]["flops"] and "fused_formulation_bytes" in fused["bwd_fused"]
    for rec in fused.values():
        assert 0.0 < rec["compulsory"]["frac"] <= rec["frac"] * 1.0000001


def test_header_is_plain_c():
    # The drop-in boundary is a C ABI: include/fvdbconv.h must compile as C11 (no torch / C++ types in the signatures).
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", str(REPO / "include" / "fvdbconv.h")], check=True)
