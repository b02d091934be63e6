"""Pins the CPU oracle: reference KATs + golden vectors produced by the reference's own oracle.

KATs restated from /root/reference tests/unit/test_conv_semantics.py:54-90,176-189 and
src/tests/GatherScatterDefaultConvTest.cu:191-255.  Golden vectors: tests/golden/make_golden.py.
"""

from itertools import product

import numpy as np
import pytest
import torch

import oracle
from oracle import Geometry


def _rows(coords):
    ijk = np.asarray(coords, dtype=np.int64).reshape(-1, 3)
    return ijk, np.zeros(ijk.shape[0], dtype=np.int64)


# ------------------------------------------------------------------ known-answer tests


def test_signed_division_and_even_torch_phase():
    assert int(oracle.floor_div(-5, 4)) == -2 and int(oracle.floor_div(5, 4)) == 1
    assert int(oracle.floor_mod(-5, 4)) == 3
    expected = {1: (0,), 2: (0, 1), 3: (-1, 0, 1), 4: (-1, 0, 1, 2), 5: (-2, -1, 0, 1, 2), 6: (-2, -1, 0, 1, 2, 3)}
    for kernel, offsets in expected.items():
        g = Geometry((kernel, 1, 1), 1)
        assert tuple(int(g.tap_offset((t, 0, 0))[0]) for t in range(kernel)) == offsets


def test_gtest_geometry_kats():
    # GatherScatterDefaultConvTest.cu:191-216
    g = Geometry((4, 3, 6), (2, 3, 4))
    assert g.semantics_version == 1 and g.kernel_volume == 72
    assert g.padding_before == (1, 1, 2) and g.padding_after == (2, 1, 3)
    assert g.tap_coord(23) == (1, 0, 5)
    assert tuple(g.tap_offset(g.tap_coord(23))) == (0, -1, 3)
    assert tuple(g.fine_from_coarse((3, -2, 1), (0, 0, 0))) == (5, -7, 2)
    g4 = Geometry(4, 4)
    coarse, ok = g4.coarse_from_fine((-2, -2, -2), (3, 3, 3))
    assert bool(ok) and tuple(coarse) == (-1, -1, -1)
    assert tuple(g4.fine_from_coarse(coarse, (3, 3, 3))) == (-2, -2, -2)
    assert not bool(g4.coarse_from_fine((-2, -2, -2), (2, 2, 2))[1])
    assert int(oracle.floor_mod(-4, 4)) == 0 and int(oracle.floor_mod(-3, 4)) != 0
    # :218-256 direct projection is phase aware for K = S = 1..6 and mixed (2,3,4)
    for k in range(1, 7):
        gk = Geometry(k, k)
        for fine in [(-7, -5, -3), (-2, -1, 0), (0, 1, 2), (3, 4, 5), (8, 7, 6)]:
            shifted = np.asarray(fine) + np.asarray(gk.padding_before)
            coarse, tap = shifted // k, shifted % k
            rebuilt, ok = gk.coarse_from_fine(fine, tap)
            assert bool(ok) and tuple(rebuilt) == tuple(coarse)
            assert tuple(gk.fine_from_coarse(coarse, tap)) == fine
    mixed = Geometry((2, 3, 4), (2, 3, 4))
    shifted = np.asarray((-3, -4, 6)) + np.asarray(mixed.padding_before)
    assert tuple(mixed.fine_from_coarse(shifted // np.array([2, 3, 4]), shifted % np.array([2, 3, 4]))) == (-3, -4, 6)
    assert Geometry(2, 2).padding_before == (0, 0, 0)
    with pytest.raises(ValueError):
        Geometry((0, 1, 1), 1)
    with pytest.raises(ValueError):
        Geometry(3, (1, -1, 1))


def test_issue_668_endpoint_counts():
    fine = [(c, 0, 0) for c in range(16)]
    assert oracle.forward_degrees(fine, (4, 1, 1), (4, 1, 1)) == {(0, 0, 0): 3, (1, 0, 0): 4, (2, 0, 0): 4, (3, 0, 0): 4, (4, 0, 0): 1}


def test_issue_668_16_cubed_round_trip():
    fine = list(product(range(16), repeat=3))
    ijk, b = _rows(fine)
    coarse, _ = oracle.conv_grid(ijk, b, 4, 4)
    assert coarse.shape[0] == 5**3
    back, _ = oracle.conv_transpose_grid(coarse, np.zeros(len(coarse), dtype=np.int64), 4, 4)
    assert set(fine).issubset({tuple(r) for r in back.tolist()})


@pytest.mark.parametrize("kernel", range(1, 7))
def test_kernel_equals_stride_projection(kernel):
    g = Geometry(kernel, kernel)
    for fine in product((-kernel, -1, 0, kernel - 1, kernel), repeat=3):
        edges = oracle.relation_edges([fine], kernel, kernel)
        assert len(edges) == 1
        _, coarse, tap = edges[0]
        assert coarse == tuple((fine[a] + g.padding_before[a]) // kernel for a in range(3))
        assert tuple(g.fine_from_coarse(coarse, tap)) == fine


def test_transpose_kat_19_43():
    # tests/unit/test_conv_semantics.py:176-189
    coarse = [(-1, 0, 0), (2, 0, 0)]
    features = torch.tensor([[5.0, 7.0], [11.0, 13.0]], dtype=torch.float64)
    weights = torch.zeros((2, 2, 2, 1, 1), dtype=torch.float64)
    weights[:, :, 0, 0, 0] = torch.tensor([[1.0, 2.0], [3.0, 4.0]])
    weights[:, :, 1, 0, 0] = torch.tensor([[5.0, 6.0], [7.0, 8.0]])
    fine, fb = oracle.conv_transpose_grid(*_rows(coarse), (2, 1, 1), (2, 1, 1))
    topo = oracle.build_topology(*_rows(coarse), fine, fb, (2, 1, 1), (2, 1, 1), transposed=True)
    out = oracle.gs_conv(features, weights, topo)
    row = [tuple(r) for r in fine.tolist()].index((-2, 0, 0))
    assert out[row].tolist() == [19.0, 43.0]
    assert {tuple(r) for r in fine.tolist()} == oracle.transpose_support(coarse, (2, 1, 1), (2, 1, 1))


# ------------------------------------------------------------------ golden vectors from the reference oracle


def test_semantics_golden(semantics_golden):
    assert len(semantics_golden) > 100
    for case in semantics_golden:
        ks, st = tuple(case["kernel_size"]), tuple(case["stride"])
        g = Geometry(ks, st)
        assert list(g.padding_before) == case["p_before"] and list(g.padding_after) == case["p_after"]
        assert [int(g.tap_offset((t, 0, 0))[0]) for t in range(ks[0])] == case["offsets_axis0"]
        fine = [tuple(c) for c in case["fine"]]
        edges = oracle.relation_edges(fine, ks, st)
        assert [[list(f), list(c), list(t)] for f, c, t in edges] == case["edges"]
        assert sorted(oracle.forward_degrees(fine, ks, st).items()) == [(tuple(c), d) for c, d in case["forward_degrees"]]
        ijk, b = _rows(fine)
        coarse, cb = oracle.conv_grid(ijk, b, ks, st)
        assert coarse.tolist() == case["forward_support"]
        back, _ = oracle.conv_transpose_grid(coarse, cb, ks, st)
        assert back.tolist() == case["transpose_support_of_forward"]
        # CSR map == relation edges (keyed by coordinate), forward and transposed builders
        topo = oracle.build_topology(ijk, b, coarse, cb, ks, st)
        got = {(tuple(ijk[gi]), tuple(coarse[so]), g.tap_coord(tap)) for tap, gi, so in oracle.topology_edge_set(topo.gather_indices, topo.scatter_indices, topo.offsets)}
        want = {(tuple(f), tuple(c), tuple(t)) for f, c, t in case["edges"]}
        assert got == want
        topo_t = oracle.build_topology(coarse, cb, ijk, b, ks, st, transposed=True)
        got_t = {(tuple(ijk[so]), tuple(coarse[gi]), g.tap_coord(tap)) for tap, gi, so in oracle.topology_edge_set(topo_t.gather_indices, topo_t.scatter_indices, topo_t.offsets)}
        assert got_t == want
        rev = oracle.reverse_topology(topo)
        assert rev.gather_indices is topo.scatter_indices and rev.is_transposed
        assert oracle.topology_edge_set(rev.scatter_indices, rev.gather_indices, rev.offsets) == oracle.topology_edge_set(topo_t.scatter_indices, topo_t.gather_indices, topo_t.offsets)


def test_dense_golden_values_and_gradients(dense_golden):
    data, meta = dense_golden
    assert len(meta) == 12
    for case in meta:
        key, ks, st, transposed = case["key"], case["kernel_size"], case["stride"], case["transposed"]
        src, tgt = data[key + "_source"], data[key + "_target"]
        features = torch.from_numpy(data[key + "_features"]).requires_grad_()
        weights = torch.from_numpy(data[key + "_weights"]).requires_grad_()
        sb, tb = np.zeros(len(src), dtype=np.int64), np.zeros(len(tgt), dtype=np.int64)
        gen, gb = (oracle.conv_transpose_grid if transposed else oracle.conv_grid)(src, sb, ks, st)
        assert gen.tolist() == tgt.tolist()  # generated target == reference support (both sorted)
        topo = oracle.build_topology(src, sb, tgt, tb, ks, st, transposed=transposed)
        values = oracle.gs_conv(features.detach(), weights.detach(), topo)
        torch.testing.assert_close(values, torch.from_numpy(data[key + "_values"]), rtol=1e-11, atol=1e-11)
        probe = torch.arange(1, values.numel() + 1, dtype=torch.float64).reshape_as(values)
        grad_f, grad_w = oracle.gs_conv_backward(probe, features.detach(), weights.detach(), topo)
        torch.testing.assert_close(grad_f, torch.from_numpy(data[key + "_grad_features"]), rtol=1e-11, atol=1e-11)
        torch.testing.assert_close(grad_w, torch.from_numpy(data[key + "_grad_weights"]), rtol=1e-11, atol=1e-11)
        # the oracle's own dense restatement agrees too
        fn = oracle.dense_transpose_oracle if transposed else oracle.dense_forward_oracle
        dense, origin = fn(src.tolist(), features.detach(), weights.detach(), ks, st)
        local = tgt.astype(np.int64) - np.asarray(origin)
        torch.testing.assert_close(dense[0][:, local[:, 0], local[:, 1], local[:, 2]].t(), values, rtol=1e-11, atol=1e-11)


# ------------------------------------------------------------------ oracle self-consistency on random inputs


def _random_batch(seed, n=400, extent=12, batches=3):
    rng = np.random.default_rng(seed)
    ijk = rng.integers(-extent, extent, size=(n, 3))
    b = rng.integers(0, batches, size=n)
    table = np.unique(np.concatenate([b[:, None], ijk], axis=1), axis=0)
    perm = rng.permutation(len(table))
    return table[perm, 1:], table[perm, 0]


@pytest.mark.parametrize("ks,st", [(3, 1), (2, 2), ((3, 5, 1), (1, 2, 3)), (4, 1), (3, 2)])
def test_adjoint_identity_and_flip(ks, st):
    # adjoint <y, d> == <x, L^T d> (GatherScatterDefaultConvTest.cu:1175-1201)
    ijk, b = _random_batch(50)
    out, ob = oracle.conv_grid(ijk, b, ks, st)
    topo = oracle.build_topology(ijk, b, out, ob, ks, st)
    g = Geometry(ks, st)
    gen = torch.Generator().manual_seed(50)
    x = torch.randn((len(ijk), 3), generator=gen, dtype=torch.float64)
    w = torch.randn((4, 3, *g.kernel_size), generator=gen, dtype=torch.float64)
    d = torch.randn((len(out), 4), generator=gen, dtype=torch.float64)
    y = oracle.gs_conv(x, w, topo)
    lt_d = oracle.gs_conv(d, w.transpose(0, 1).contiguous(), oracle.reverse_topology(topo))
    torch.testing.assert_close((y * d).sum(), (x * lt_d).sum(), rtol=1e-12, atol=1e-10)
    grad_x, grad_w = oracle.gs_conv_backward(d, x, w, topo)
    torch.testing.assert_close(grad_x, lt_d, rtol=1e-12, atol=1e-12)
    xa, wa = x.clone().requires_grad_(), w.clone().requires_grad_()
    # autograd through an index_select formulation
    wp = oracle.permute_weights(wa)
    ya = torch.zeros_like(y)
    gi = torch.from_numpy(topo.gather_indices.astype(np.int64))
    si = torch.from_numpy(topo.scatter_indices.astype(np.int64))
    for k in range(topo.kernel_volume):
        s, e = int(topo.offsets[k]), int(topo.offsets[k + 1])
        ya = ya.index_add(0, si[s:e], xa[gi[s:e]] @ wp[k])
    ga = torch.autograd.grad((ya * d).sum(), (xa, wa))
    torch.testing.assert_close(grad_x, ga[0], rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(grad_w, ga[1], rtol=1e-12, atol=1e-12)


def test_flip_identity_stride1():
    # convT(x, W) == conv(x, flip W) at stride 1, odd K (GatherScatterDefaultConvTest.cu:1389-1457)
    ijk, b = _random_batch(70, batches=1)
    gen = torch.Generator().manual_seed(70)
    x = torch.randn((len(ijk), 2), generator=gen, dtype=torch.float64)
    w = torch.randn((3, 2, 3, 3, 3), generator=gen, dtype=torch.float64)
    fwd = oracle.build_topology(ijk, b, ijk, b, 3, 1)
    tr = oracle.build_topology(ijk, b, ijk, b, 3, 1, transposed=True)
    torch.testing.assert_close(oracle.gs_conv(x, w, tr), oracle.gs_conv(x, w.flip(2, 3, 4), fwd), rtol=1e-12, atol=1e-12)


def test_same_grid_dense_map_symmetry():
    ijk, b = _random_batch(3)
    nbr = oracle.dense_kernel_map(ijk, b, ijk, b, 3, 1)
    valid = np.argwhere(nbr >= 0)
    for o, k in valid[:500]:
        assert nbr[nbr[o, k], 26 - k] == o
    assert (nbr[:, 13] == np.arange(len(ijk))).all()


def test_neighbor_indexes_and_row_order():
    ijk, b = _random_batch(9)
    order = oracle.index_grid_row_order(b, ijk)
    ijk, b = ijk[order], b[order]
    assert (np.diff(b) >= 0).all()
    offsets = np.concatenate([[0], np.cumsum(np.bincount(b, minlength=3))])
    nbr = oracle.neighbor_indexes(ijk, b, offsets, ijk, b, 1)
    assert nbr.shape == (len(ijk), 3, 3, 3)
    centre = nbr[:, 1, 1, 1]
    assert (centre == np.arange(len(ijk)) - offsets[b]).all()
    dense = oracle.dense_kernel_map(ijk, b, ijk, b, 3, 1)
    local = np.where(dense >= 0, dense - offsets[b][:, None], -1)
    assert (nbr.reshape(len(ijk), 27) == local).all()


def test_half_promotion_and_empty():
    ijk, b = _random_batch(5, n=50)
    topo = oracle.build_topology(ijk, b, ijk, b, 3, 1)
    x = torch.randn(len(ijk), 4, dtype=torch.bfloat16)
    w = torch.randn(4, 4, 3, 3, 3, dtype=torch.float32)
    assert oracle.gs_conv(x, w, topo).dtype == torch.float32  # result_type promotion (:850-852)
    assert oracle.gs_conv(x, w.bfloat16(), topo).dtype == torch.bfloat16
    empty = oracle.build_topology(np.zeros((0, 3)), np.zeros(0), np.zeros((0, 3)), np.zeros(0), 3, 1)
    assert empty.total_pairs == 0 and empty.offsets.tolist() == [0] * 28
    assert oracle.gs_conv(torch.zeros(0, 4), w, empty).shape == (0, 4)
    gf, gw = oracle.gs_conv_backward(torch.zeros(0, 4), torch.zeros(0, 4), w, empty)
    assert gf.shape == (0, 4) and gw.shape == w.shape and not gw.any()


# ------------------------------------------------------------------ pooling / refinement oracle


def test_pool_oracle_matches_dense_torch_pooling():
    # On a dense block the sparse definitions must equal torch's dense pooling / nearest upsampling (what the reference's own
    # tests compare against); this pins oracle/pool_oracle.py.
    from oracle import pool_oracle as po

    rng = np.random.default_rng(0)
    dim, c = 8, 5
    ijk = np.stack(np.meshgrid(np.arange(dim), np.arange(dim), np.arange(dim), indexing="ij"), -1).reshape(-1, 3)
    bidx = np.zeros(len(ijk), dtype=np.int64)
    x = rng.standard_normal((len(ijk), c))
    dense = torch.from_numpy(x.reshape(dim, dim, dim, c)).permute(3, 0, 1, 2)[None]
    for factor in ((2, 2, 2), (2, 4, 1)):
        coarse = po.coarsened_ijk(ijk, factor)
        assert len(coarse) == (dim // factor[0]) * (dim // factor[1]) * (dim // factor[2])
        cb = np.zeros(len(coarse), dtype=np.int64)
        for mode, fn in (("max", torch.nn.functional.max_pool3d), ("avg", torch.nn.functional.avg_pool3d)):
            y, children = po.pool(ijk, bidx, x, coarse, cb, factor, (0, 0, 0), mode)
            want = fn(dense, kernel_size=factor)[0].permute(1, 2, 3, 0).numpy()
            got = np.zeros_like(want)
            got[coarse[:, 0], coarse[:, 1], coarse[:, 2]] = y
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
            assert (children >= 0).all()
        up, parent = po.refine(coarse, cb, y, ijk, bidx, factor)
        np.testing.assert_array_equal(up, y[parent])
        dxs = po.refine_backward(np.ones_like(up), parent, len(coarse))
        assert (dxs == factor[0] * factor[1] * factor[2]).all()
        fine_again = po.refined_ijk(coarse, factor)
        assert sorted(map(tuple, fine_again.tolist())) == sorted(map(tuple, ijk.tolist()))
    # metadata: block-centroid transforms are mutually inverse
    s, o = po.coarse_metadata((0.5, 1.0, 2.0), (3.0, -2.0, 7.0), (2, 3, 4))
    s2, o2 = po.fine_metadata(s, o, (2, 3, 4))
    np.testing.assert_allclose(s2, (0.5, 1.0, 2.0)), np.testing.assert_allclose(o2, (3.0, -2.0, 7.0))
    # max backward on a sparse signed set: the gradient goes to the first maximal child only
    pts = np.array([[-1, 0, 0], [-2, 0, 0], [0, 0, 0], [5, 5, 5]])
    xv = np.array([[1.0], [3.0], [2.0], [7.0]])
    coarse = po.coarsened_ijk(pts, (2, 2, 2))
    assert sorted(map(tuple, coarse.tolist())) == [(-1, 0, 0), (0, 0, 0), (2, 2, 2)]
    y, children = po.pool(pts, np.zeros(4, dtype=np.int64), xv, coarse, np.zeros(3, dtype=np.int64), (2, 2, 2), (0, 0, 0), "max")
    assert y[:, 0].tolist() == [3.0, 2.0, 7.0]
    dx = po.pool_backward(np.array([[10.0], [20.0], [30.0]]), xv, children, 4, "max")
    assert dx[:, 0].tolist() == [0.0, 10.0, 20.0, 30.0]
