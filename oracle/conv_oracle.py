"""CPU restatement of the reference ConvolutionPlan hot path (TEST INFRASTRUCTURE, see __init__).

Integer work (geometry, kernel maps, generated topologies, neighbour indices) is numpy; the
floating-point executor is torch-on-CPU because the reference's own CPU path is ``torch::mm``
(``src/fvdb/detail/ops/convolution/GatherScatterDefault.cu:605-615``).  Every function cites the
reference lines it follows (paths relative to the reference checkout).

Row numbering: the reference never pins voxel row order (every test keys by coordinate), so all
functions here take coordinate tables *in the caller's row order* and answer in those row numbers.
"""

from __future__ import annotations

from dataclasses import dataclass
from itertools import product
from typing import Iterable, Sequence

import numpy as np
import torch

__all__ = [
    "Geometry",
    "Topology",
    "floor_div",
    "floor_mod",
    "normalize_3d",
    "CoordIndex",
    "build_topology",
    "reverse_topology",
    "dense_kernel_map",
    "topology_edge_set",
    "conv_grid",
    "conv_transpose_grid",
    "index_grid_row_order",
    "neighbor_indexes",
    "permute_weights",
    "gs_conv",
    "gs_conv_backward",
    "relation_edges",
    "forward_degrees",
    "forward_support",
    "transpose_support",
    "dense_forward_oracle",
    "dense_transpose_oracle",
]


# ---------------------------------------------------------------------------------------------
# Geometry -- src/fvdb/detail/ops/convolution/ConvolutionGeometry.h:30-207
# ---------------------------------------------------------------------------------------------


def normalize_3d(value) -> tuple[int, int, int]:
    if isinstance(value, (int, np.integer)):
        return (int(value),) * 3
    value = [int(v) for v in (value.tolist() if hasattr(value, "tolist") else value)]
    if len(value) == 1:
        return (value[0],) * 3
    if len(value) != 3:
        raise ValueError(f"expected three values, got {value}")
    return tuple(value)  # type: ignore[return-value]


def floor_div(dividend, divisor):
    """Euclidean floor division for a positive divisor (ConvolutionGeometry.h:127-135)."""
    return np.floor_divide(dividend, divisor)


def floor_mod(dividend, divisor):
    """Euclidean modulo in [0, divisor) (ConvolutionGeometry.h:138-142)."""
    return np.mod(dividend, divisor)


class Geometry:
    """``fine = stride * coarse + tap - padding_before`` (ConvolutionGeometry.h:21-38, 93-104)."""

    semantics_version = 1  # ConvolutionGeometry.h:32
    phase_policy = "torch_same_phase"  # src/python/Bindings.cpp:539-540

    def __init__(self, kernel_size, stride):
        self.kernel_size = normalize_3d(kernel_size)
        self.stride = normalize_3d(stride)
        for d in range(3):  # ConvolutionGeometry.h:150-172
            if self.kernel_size[d] <= 0:
                raise ValueError(f"kernel_size must be strictly positive, got {self.kernel_size[d]} in dimension {d}")
            if self.stride[d] <= 0:
                raise ValueError(f"stride must be strictly positive, got {self.stride[d]} in dimension {d}")
        # ConvolutionGeometry.h:174-183
        self.padding_before = tuple((k - 1) // 2 for k in self.kernel_size)
        self.padding_after = tuple(k - 1 - p for k, p in zip(self.kernel_size, self.padding_before))
        self.kernel_volume = self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        self.dilation = (1, 1, 1)  # ConvolutionGeometry.h:73-76
        self.registration_offset = (0, 0, 0)  # ConvolutionGeometry.h:78-81

    def tap_coord(self, tap_index: int) -> tuple[int, int, int]:
        """x-major linear tap index -> zero-based tap coordinate (ConvolutionGeometry.h:85-90)."""
        k0, k1, k2 = self.kernel_size
        return (tap_index // (k1 * k2), (tap_index // k2) % k1, tap_index % k2)

    def taps(self) -> np.ndarray:
        """All tap coordinates, ``[K^3, 3]`` int64, in linear-index order."""
        return np.array([self.tap_coord(k) for k in range(self.kernel_volume)], dtype=np.int64).reshape(-1, 3)

    def tap_offset(self, tap) -> np.ndarray:
        return np.asarray(tap, dtype=np.int64) - np.asarray(self.padding_before, dtype=np.int64)

    def fine_from_coarse(self, coarse, tap) -> np.ndarray:
        """ConvolutionGeometry.h:99-104."""
        coarse = np.asarray(coarse, dtype=np.int64)
        return coarse * np.asarray(self.stride, dtype=np.int64) + self.tap_offset(tap)

    def coarse_from_fine(self, fine, tap):
        """Solve the relation for coarse when every axis divides (ConvolutionGeometry.h:107-124).

        Returns ``(coarse, ok)``; rows with ``ok == False`` hold the floor quotient and must be ignored.
        """
        fine = np.asarray(fine, dtype=np.int64)
        numer = fine - self.tap_offset(tap)
        stride = np.asarray(self.stride, dtype=np.int64)
        ok = np.all(floor_mod(numer, stride) == 0, axis=-1)
        return floor_div(numer, stride), ok


# ---------------------------------------------------------------------------------------------
# Coordinate index (stands in for the NanoVDB accessor: isActive / getValue-1)
# ---------------------------------------------------------------------------------------------


class CoordIndex:
    """Exact ``(batch, i, j, k) -> row`` lookup over a coordinate table given in row order.

    Plays the role of ``feat_grid->getAccessor()`` + ``voxelOffset(batch)`` in
    GatherScatterDefault.cu:126-127,186-201: ``lookup`` returns the batch-cumulative row or -1.
    """

    def __init__(self, ijk: np.ndarray, bidx: np.ndarray):
        ijk = np.asarray(ijk, dtype=np.int64).reshape(-1, 3)
        bidx = np.asarray(bidx, dtype=np.int64).reshape(-1)
        assert ijk.shape[0] == bidx.shape[0]
        self.n = ijk.shape[0]
        table = np.concatenate([bidx[:, None], ijk], axis=1)
        # lexicographic sort on (b, i, j, k); rows are unique in a grid
        order = np.lexsort((table[:, 3], table[:, 2], table[:, 1], table[:, 0]))
        self._sorted = np.ascontiguousarray(table[order])
        self._rows = order.astype(np.int64)
        self._view = self._sorted.view([("b", np.int64), ("i", np.int64), ("j", np.int64), ("k", np.int64)]).reshape(-1)
        if self.n > 1:
            dup = np.all(self._sorted[1:] == self._sorted[:-1], axis=1)
            if dup.any():
                raise ValueError("coordinate table contains duplicate (batch, ijk) rows")
        # Fast path (full-size parity cases): when the table's bounding box fits 62 bits, (b, i, j, k) packs into ONE
        # int64 key whose order equals the lexicographic order, so a lookup is a plain integer binary search.  Queries
        # outside the box are misses by construction.  Same answers as the structured path (tests/test_oracle_golden.py).
        self._packed = None
        if self.n > 0:
            lo = self._sorted.min(axis=0)
            span = self._sorted.max(axis=0) - lo + 1
            if float(span[0]) * float(span[1]) * float(span[2]) * float(span[3]) < 2.0**62:
                self._lo, self._span = lo, span
                self._packed = self._pack(self._sorted)
        # Densest case (a few million voxels filling a box, kernel volumes of 125): a dense row table over the bounding
        # box turns a lookup into one fancy-indexing read.  Only when the box holds at most 2^26 cells.
        self._dense = None
        if self._packed is not None and int(np.prod(self._span)) <= 2**26:
            self._dense = np.full(int(np.prod(self._span)), -1, dtype=np.int32)
            self._dense[self._packed] = self._rows

    def _pack(self, table: np.ndarray) -> np.ndarray:
        rel = table - self._lo
        return ((rel[:, 0] * self._span[1] + rel[:, 1]) * self._span[2] + rel[:, 2]) * self._span[3] + rel[:, 3]

    def prepare(self, bidx: np.ndarray, ijk: np.ndarray):
        """Base queries for ``lookup_offset``: relative coordinates and packed keys (packed / dense tables only)."""
        if self._packed is None:
            return None
        ijk = np.asarray(ijk, dtype=np.int64).reshape(-1, 3)
        bidx = np.broadcast_to(np.asarray(bidx, dtype=np.int64).reshape(-1), (ijk.shape[0],))
        rel = np.concatenate([bidx[:, None], ijk], axis=1) - self._lo
        ok_b = (rel[:, 0] >= 0) & (rel[:, 0] < self._span[0])
        key = ((rel[:, 0] * self._span[1] + rel[:, 1]) * self._span[2] + rel[:, 2]) * self._span[3] + rel[:, 3]
        return rel, ok_b, key

    def lookup_offset(self, prepared, offset) -> np.ndarray:
        """``lookup(bidx, ijk + offset)`` for queries prepared once: a constant offset moves the packed key by a constant, so
        a whole stencil costs one bounds test and one table read per tap (same answers as ``lookup``)."""
        rel, inside, key = prepared
        inside = inside.copy()
        for d in range(3):
            c = rel[:, d + 1] + int(offset[d])
            inside &= (c >= 0) & (c < self._span[d + 1])
        delta = (int(offset[0]) * int(self._span[2]) + int(offset[1])) * int(self._span[3]) + int(offset[2])
        keys = key[inside] + delta
        out = np.full(rel.shape[0], -1, dtype=np.int64)
        if self._dense is not None:
            out[inside] = self._dense[keys]
            return out
        pos = np.minimum(np.searchsorted(self._packed, keys), self.n - 1)
        hit = self._packed[pos] == keys
        rows = np.full(keys.shape[0], -1, dtype=np.int64)
        rows[hit] = self._rows[pos[hit]]
        out[inside] = rows
        return out

    def lookup(self, bidx: np.ndarray, ijk: np.ndarray) -> np.ndarray:
        ijk = np.asarray(ijk, dtype=np.int64).reshape(-1, 3)
        bidx = np.broadcast_to(np.asarray(bidx, dtype=np.int64).reshape(-1), (ijk.shape[0],))
        out = np.full(ijk.shape[0], -1, dtype=np.int64)
        if self.n == 0 or ijk.shape[0] == 0:
            return out
        query = np.ascontiguousarray(np.concatenate([bidx[:, None], ijk], axis=1))
        if self._packed is not None:
            inside = np.all((query >= self._lo) & (query < self._lo + self._span), axis=1)
            keys = self._pack(query[inside])
            if self._dense is not None:
                out[inside] = self._dense[keys]
                return out
            pos = np.minimum(np.searchsorted(self._packed, keys), self.n - 1)
            hit = self._packed[pos] == keys
            rows = np.full(keys.shape[0], -1, dtype=np.int64)
            rows[hit] = self._rows[pos[hit]]
            out[inside] = rows
            return out
        qview = query.view(self._view.dtype).reshape(-1)
        pos = np.searchsorted(self._view, qview)
        pos_c = np.minimum(pos, self.n - 1)
        hit = np.all(self._sorted[pos_c] == query, axis=1) & (pos < self.n)
        out[hit] = self._rows[pos_c[hit]]
        return out


# ---------------------------------------------------------------------------------------------
# Kernel map -- GatherScatterDefault.cu:92-294
# ---------------------------------------------------------------------------------------------


@dataclass
class Topology:
    """CSR-by-tap kernel map (GatherScatterDefault.h:59-81). Index arrays int32, offsets int64."""

    gather_indices: np.ndarray
    scatter_indices: np.ndarray
    offsets: np.ndarray
    feature_total_voxels: int
    output_total_voxels: int
    kernel_volume: int
    total_pairs: int
    kernel_size: tuple[int, int, int]
    stride: tuple[int, int, int]
    is_transposed: bool


_INT32_MAX = np.iinfo(np.int32).max


def dense_kernel_map(feat_ijk, feat_bidx, out_ijk, out_bidx, kernel_size, stride, transposed=False) -> np.ndarray:
    """Output-stationary dense form of the map: ``nbr[o, k]`` = feature row or -1.

    Probe rule of GatherScatterDefault.cu:129-141: forward probes ``fineFromCoarse(out_ijk, tap)``
    on the (fine) feature grid; transposed probes ``coarseFromFine(out_ijk, tap)`` on the (coarse)
    feature grid and skips non-divisible taps.  Probes stay inside the output voxel's batch item
    (``feature_acc.grid(batch_idx)``, :126).
    """
    geometry = Geometry(kernel_size, stride)
    out_ijk = np.asarray(out_ijk, dtype=np.int64).reshape(-1, 3)
    out_bidx = np.asarray(out_bidx, dtype=np.int64).reshape(-1)
    index = CoordIndex(feat_ijk, feat_bidx)
    nbr = np.full((out_ijk.shape[0], geometry.kernel_volume), -1, dtype=np.int64)
    for k in range(geometry.kernel_volume):
        tap = geometry.tap_coord(k)
        if transposed:
            probe, ok = geometry.coarse_from_fine(out_ijk, tap)
            rows = index.lookup(out_bidx, probe)
            rows[~ok] = -1
        else:
            rows = index.lookup(out_bidx, geometry.fine_from_coarse(out_ijk, tap))
        nbr[:, k] = rows
    return nbr


def build_topology(feat_ijk, feat_bidx, out_ijk, out_bidx, kernel_size, stride, transposed=False) -> Topology:
    """Two-pass CSR build (GatherScatterDefault.cu:92-223) with the per-tap pair order canonicalised.

    The reference claims slots with racing atomics (:202-206), so pair order inside a tap segment
    is unspecified there; here each segment is ordered by output row (ascending).
    """
    geometry = Geometry(kernel_size, stride)
    n_feat = int(np.asarray(feat_ijk).reshape(-1, 3).shape[0])
    n_out = int(np.asarray(out_ijk).reshape(-1, 3).shape[0])
    if n_feat > _INT32_MAX or n_out > _INT32_MAX:  # :68-80
        raise RuntimeError("voxel count exceeds the int32 index limit")
    # one tap at a time (the dense [n_out, K] form of a 5^3 map over millions of voxels would not fit comfortably)
    K = geometry.kernel_volume
    out_ijk_ = np.asarray(out_ijk, dtype=np.int64).reshape(-1, 3)
    out_bidx_ = np.asarray(out_bidx, dtype=np.int64).reshape(-1)
    index = CoordIndex(feat_ijk, feat_bidx)
    per_tap = []
    # forward probes are fineFromCoarse(out, tap) = (S * out - pad) + tap: one prepared base query, a constant offset per tap
    prepared = None if (transposed or n_out == 0 or n_feat == 0) else index.prepare(out_bidx_, geometry.fine_from_coarse(out_ijk_, (0, 0, 0)))
    for k in range(K):
        tap = geometry.tap_coord(k)
        if prepared is not None:
            rows = index.lookup_offset(prepared, tap)
            out_rows = np.nonzero(rows >= 0)[0]
            per_tap.append((rows[out_rows].astype(np.int32), out_rows.astype(np.int32)))
            continue
        if transposed:  # :129-141
            probe, ok = geometry.coarse_from_fine(out_ijk_, tap)
            rows = index.lookup(out_bidx_, probe)
            rows[~ok] = -1
        else:
            rows = index.lookup(out_bidx_, geometry.fine_from_coarse(out_ijk_, tap))
        out_rows = np.nonzero(rows >= 0)[0]
        per_tap.append((rows[out_rows].astype(np.int32), out_rows.astype(np.int32)))
    offsets = np.zeros(K + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([len(g) for g, _ in per_tap])  # :145-151
    total = int(offsets[K])
    gather = np.concatenate([g for g, _ in per_tap]) if total else np.empty(0, dtype=np.int32)
    scatter = np.concatenate([s for _, s in per_tap]) if total else np.empty(0, dtype=np.int32)
    return Topology(gather, scatter, offsets, n_feat, n_out, K, total, geometry.kernel_size, geometry.stride, bool(transposed))


def reverse_topology(topology: Topology) -> Topology:
    """Alias-swap view (GatherScatterDefault.cu:273-294): no copy, direction flipped."""
    return Topology(
        topology.scatter_indices,
        topology.gather_indices,
        topology.offsets,
        topology.output_total_voxels,
        topology.feature_total_voxels,
        topology.kernel_volume,
        topology.total_pairs,
        topology.kernel_size,
        topology.stride,
        not topology.is_transposed,
    )


def topology_edge_set(gather, scatter, offsets) -> set[tuple[int, int, int]]:
    """Order-independent ``(tap, feature_row, output_row)`` set (tests/unit/test_conv_semantics_integration.py:60-69)."""
    gather = np.asarray(gather).astype(np.int64)
    scatter = np.asarray(scatter).astype(np.int64)
    offsets = np.asarray(offsets).astype(np.int64)
    taps = np.repeat(np.arange(len(offsets) - 1, dtype=np.int64), np.diff(offsets))
    return set(zip(taps.tolist(), gather.tolist(), scatter.tolist()))


# ---------------------------------------------------------------------------------------------
# Generated target topologies -- BuildGridForConv.cu:465-544, BuildGridForConvTranspose.cu:310-359
# ---------------------------------------------------------------------------------------------


def _unique_sorted(bidx: np.ndarray, ijk: np.ndarray):
    if ijk.shape[0] == 0:
        return np.zeros((0, 3), dtype=np.int64), np.zeros((0,), dtype=np.int64)
    table = np.unique(np.concatenate([bidx[:, None], ijk], axis=1), axis=0)
    return table[:, 1:].copy(), table[:, 0].copy()


def conv_grid(ijk, bidx, kernel_size, stride):
    """Complete forward support: every divisible tap of every fine voxel emits a coarse voxel.

    BuildGridForConv.cu:465-525 (CPU path = the plain definition); ``K == S`` reduces to
    ``floorDiv(fine + padBefore, S)`` (:485-499).  Returns ``(ijk, bidx)`` lexicographically
    sorted by ``(batch, i, j, k)`` -- a *set*; the product's row order is its own.
    """
    geometry = Geometry(kernel_size, stride)
    ijk = np.asarray(ijk, dtype=np.int64).reshape(-1, 3)
    bidx = np.asarray(bidx, dtype=np.int64).reshape(-1)
    coords, batches = [], []
    for k in range(geometry.kernel_volume):
        coarse, ok = geometry.coarse_from_fine(ijk, geometry.tap_coord(k))
        coords.append(coarse[ok])
        batches.append(bidx[ok])
    return _unique_sorted(np.concatenate(batches), np.concatenate(coords))


def conv_transpose_grid(ijk, bidx, kernel_size, stride):
    """Complete transposed support: every coarse voxel spreads through every tap.

    BuildGridForConvTranspose.cu:310-342.
    """
    geometry = Geometry(kernel_size, stride)
    ijk = np.asarray(ijk, dtype=np.int64).reshape(-1, 3)
    bidx = np.asarray(bidx, dtype=np.int64).reshape(-1)
    coords = [geometry.fine_from_coarse(ijk, geometry.tap_coord(k)) for k in range(geometry.kernel_volume)]
    return _unique_sorted(np.tile(bidx, geometry.kernel_volume), np.concatenate(coords))


def index_grid_row_order(bidx, ijk) -> np.ndarray:
    """Permutation that sorts voxels into the index-grid row order this build adopts.

    Upstream NanoVDB convention recalled in SURVEY.md section 8c (unverifiable offline, so parity is
    still defined through ijk): rows ordered by batch, then root tile ``(x>>12, y>>12, z>>12)``
    (signed, x-major), then upper-node child offset ``((x>>7)&31, (y>>7)&31, (z>>7)&31)``, lower
    ``((x>>3)&15, ...)`` and leaf ``(x&7, y&7, z&7)``, each x-major.
    """
    ijk = np.asarray(ijk, dtype=np.int64).reshape(-1, 3)
    bidx = np.asarray(bidx, dtype=np.int64).reshape(-1)
    keys = []
    for shift, mask in ((0, 7), (3, 15), (7, 31)):
        for axis in (2, 1, 0):
            keys.append((ijk[:, axis] >> shift) & mask)
    for axis in (2, 1, 0):
        keys.append(ijk[:, axis] >> 12)
    keys.append(bidx)
    return np.lexsort(tuple(keys))


# ---------------------------------------------------------------------------------------------
# Neighbour indices -- src/fvdb/detail/ops/NeighborIndexes.cu:22-47
# ---------------------------------------------------------------------------------------------


def neighbor_indexes(grid_ijk, grid_bidx, voxel_offsets, query_ijk, query_bidx, extent: int, shift: int = 0) -> np.ndarray:
    """Per-grid-local neighbour index or -1 over ``[-extent, extent]^3`` (x-major), int64.

    NeighborIndexes.cu:35-44: ``ijk0 = query << shift``; value is ``getValue - 1`` inside the query's
    own grid, i.e. the batch-cumulative row minus that grid's voxel offset.
    """
    index = CoordIndex(grid_ijk, grid_bidx)
    query_ijk = np.asarray(query_ijk, dtype=np.int64).reshape(-1, 3)
    query_bidx = np.asarray(query_bidx, dtype=np.int64).reshape(-1)
    voxel_offsets = np.asarray(voxel_offsets, dtype=np.int64)
    width = 2 * extent + 1
    out = np.full((query_ijk.shape[0], width, width, width), -1, dtype=np.int64)
    base = query_ijk << shift
    for a, b, c in product(range(width), repeat=3):
        rows = index.lookup(query_bidx, base + np.array([a - extent, b - extent, c - extent], dtype=np.int64))
        local = rows - voxel_offsets[query_bidx]
        out[:, a, b, c] = np.where(rows >= 0, local, -1)
    return out


# ---------------------------------------------------------------------------------------------
# Executor -- GatherScatterDefault.cu:592-924
# ---------------------------------------------------------------------------------------------


def permute_weights(weights: torch.Tensor) -> torch.Tensor:
    """``[Cout, Cin, k0, k1, k2] -> [K^3, Cin, Cout]`` contiguous (GatherScatterDefault.cu:691)."""
    k3 = weights.shape[2] * weights.shape[3] * weights.shape[4]
    return weights.permute(2, 3, 4, 1, 0).reshape(k3, weights.shape[1], weights.shape[0]).contiguous()


def _mm_safe(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """CPU half/bfloat16 matmul promotes to fp32 and demotes (GatherScatterDefault.cu:605-615)."""
    if a.dtype in (torch.float16, torch.bfloat16):
        return torch.mm(a.float(), b.float()).to(a.dtype)
    return torch.mm(a, b)


def _check_conv(features: torch.Tensor, weights: torch.Tensor, topology: Topology, name: str) -> None:
    """GatherScatterDefault.cu:635-667."""
    if features.dim() != 2:
        raise RuntimeError(f"{name}: features must be 2D")
    if features.shape[0] != topology.feature_total_voxels:
        raise RuntimeError(f"{name}: features.size(0)={features.shape[0]} must match featureTotalVoxels={topology.feature_total_voxels}")
    if not features.is_floating_point() or not weights.is_floating_point():
        raise RuntimeError(f"{name}: features and weights must be floating point")
    if weights.dim() != 5:
        raise RuntimeError(f"{name}: weights must be 5D [C_out, C_in, k0, k1, k2]")
    if features.shape[1] != weights.shape[1]:
        raise RuntimeError(f"{name}: features channels={features.shape[1]} must match weights C_in={weights.shape[1]}")
    if tuple(weights.shape[2:]) != tuple(topology.kernel_size):
        raise RuntimeError(f"{name}: weights spatial dims must match topology kernel_size")


def gs_conv(features: torch.Tensor, weights: torch.Tensor, topology: Topology, *, accumulate_dtype=None) -> torch.Tensor:
    """Forward (either direction; the arrays are already oriented): per tap gather -> mm -> scatter-add.

    GatherScatterDefault.cu:673-724 with the promotion of :850-852.  ``accumulate_dtype`` is an
    oracle-only knob: ``None`` follows the reference (accumulate in the promoted working dtype, so
    bf16 rounds after every tap); ``torch.float32`` gives the fp32 reference the north-star
    tolerance for half types is stated against.
    """
    _check_conv(features, weights, topology, "gs_conv")
    working = torch.result_type(features, weights)
    if accumulate_dtype is not None:
        working = accumulate_dtype
    features = features.to(working)
    w = permute_weights(weights).to(working)
    out = torch.zeros((topology.output_total_voxels, weights.shape[0]), dtype=working)
    if topology.output_total_voxels == 0 or topology.kernel_volume == 0 or topology.total_pairs == 0:
        return out  # :698-699
    gather = torch.from_numpy(topology.gather_indices.astype(np.int64))
    scatter = torch.from_numpy(topology.scatter_indices.astype(np.int64))
    for k in range(topology.kernel_volume):  # :706-721
        start, end = int(topology.offsets[k]), int(topology.offsets[k + 1])
        if end == start:
            continue
        a_k = features.index_select(0, gather[start:end])
        out.index_add_(0, scatter[start:end], _mm_safe(a_k, w[k]))
    return out


def gs_conv_backward(grad_output, features, weights, topology: Topology, *, accumulate_dtype=None):
    """dgrad + wgrad (GatherScatterDefault.cu:742-816); returns ``(grad_features, grad_weights)``.

    ``grad_weights`` comes back contiguous in the public ``[Cout, Cin, k0, k1, k2]`` layout (:810-813).
    """
    _check_conv(features, weights, topology, "gs_conv_backward")
    if grad_output.dim() != 2 or grad_output.shape[0] != topology.output_total_voxels:
        raise RuntimeError("grad_output shape mismatch")  # :869-870
    working = torch.result_type(features, weights)
    if accumulate_dtype is not None:
        working = accumulate_dtype
    features = features.to(working)
    grad_output = grad_output.to(working)
    c_out, c_in = weights.shape[0], weights.shape[1]
    w = permute_weights(weights).to(working)
    grad_features = torch.zeros((topology.feature_total_voxels, c_in), dtype=working)
    grad_w = torch.zeros((topology.kernel_volume, c_in, c_out), dtype=working)
    if not (topology.output_total_voxels == 0 or topology.kernel_volume == 0 or topology.total_pairs == 0):
        gather = torch.from_numpy(topology.gather_indices.astype(np.int64))
        scatter = torch.from_numpy(topology.scatter_indices.astype(np.int64))
        for k in range(topology.kernel_volume):  # :786-808
            start, end = int(topology.offsets[k]), int(topology.offsets[k + 1])
            if end == start:
                continue
            feat_buf = features.index_select(0, gather[start:end])
            grad_buf = grad_output.index_select(0, scatter[start:end])
            grad_features.index_add_(0, gather[start:end], _mm_safe(grad_buf, w[k].t()))
            grad_w[k] = _mm_safe(feat_buf.t(), grad_buf)
    k0, k1, k2 = topology.kernel_size
    grad_weights = grad_w.reshape(k0, k1, k2, c_in, c_out).permute(4, 3, 0, 1, 2).contiguous()
    return grad_features, grad_weights


# ---------------------------------------------------------------------------------------------
# Independent scalar / dense restatement of fvdb/utils/tests/convolution_semantics_oracle.py.
# Used to cross-check the restatement above against torch's dense conv3d on small cases and,
# through tests/golden, against the reference's own copy of these functions.
# ---------------------------------------------------------------------------------------------

Coord = tuple[int, int, int]


def _as_coords(coordinates: Iterable[Sequence[int]]) -> list[Coord]:
    return [tuple(int(c) for c in coordinate) for coordinate in coordinates]  # type: ignore[misc]


def relation_edges(fine_coordinates, kernel_size, stride, coarse_coordinates=None) -> list[tuple[Coord, Coord, Coord]]:
    """Sorted ``(fine, coarse, tap)`` edges of the canonical relation (oracle :131-147)."""
    geometry = Geometry(kernel_size, stride)
    fine_set = set(_as_coords(fine_coordinates))
    coarse_set = None if coarse_coordinates is None else set(_as_coords(coarse_coordinates))
    edges = set()
    for fine in fine_set:
        for k in range(geometry.kernel_volume):
            tap = geometry.tap_coord(k)
            coarse, ok = geometry.coarse_from_fine(np.array(fine), tap)
            if bool(ok):
                coarse_t = tuple(int(c) for c in coarse)
                if coarse_set is None or coarse_t in coarse_set:
                    edges.add((fine, coarse_t, tap))
    return sorted(edges)


def forward_degrees(fine_coordinates, kernel_size, stride) -> dict[Coord, int]:
    degrees: dict[Coord, int] = {}
    for _, coarse, _ in relation_edges(fine_coordinates, kernel_size, stride):
        degrees[coarse] = degrees.get(coarse, 0) + 1
    return degrees


def forward_support(fine_coordinates, kernel_size, stride) -> set[Coord]:
    return set(forward_degrees(fine_coordinates, kernel_size, stride))


def transpose_support(coarse_coordinates, kernel_size, stride) -> set[Coord]:
    geometry = Geometry(kernel_size, stride)
    support = set()
    for coarse in set(_as_coords(coarse_coordinates)):
        for k in range(geometry.kernel_volume):
            support.add(tuple(int(c) for c in geometry.fine_from_coarse(np.array(coarse), geometry.tap_coord(k))))
    return support


def dense_forward_oracle(fine_coordinates, features: torch.Tensor, weights: torch.Tensor, kernel_size, stride):
    """torch ``conv3d`` (padding 0) on a global-coordinate canvas (oracle :199-235).

    Returns ``(values [1, Cout, X, Y, Z], origin)``; ``values[0, :, c - origin]`` is the output at
    coarse coordinate ``c``.
    """
    geometry = Geometry(kernel_size, stride)
    fine = np.asarray(_as_coords(fine_coordinates), dtype=np.int64).reshape(-1, 3)
    stride_v = np.asarray(geometry.stride)
    r_min = -np.asarray(geometry.padding_before)
    r_max = np.asarray(geometry.kernel_size) - 1 - np.asarray(geometry.padding_before)
    coarse_min = np.floor_divide(fine.min(axis=0) - r_max, stride_v)
    coarse_max = -np.floor_divide(-(fine.max(axis=0) - r_min), stride_v)
    input_min = geometry.fine_from_coarse(coarse_min, (0, 0, 0))
    input_max = geometry.fine_from_coarse(coarse_max, tuple(k - 1 for k in geometry.kernel_size))
    shape = tuple(int(v) for v in (input_max - input_min + 1))
    dense = torch.zeros((1, features.shape[1], *shape), dtype=features.dtype)
    local = fine - input_min
    dense[0, :, local[:, 0], local[:, 1], local[:, 2]] = features.t()
    values = torch.nn.functional.conv3d(dense, weights, stride=geometry.stride, padding=0)
    return values, tuple(int(v) for v in coarse_min)


def dense_transpose_oracle(coarse_coordinates, features: torch.Tensor, weights: torch.Tensor, kernel_size, stride):
    """torch ``conv_transpose3d`` with ``weights.transpose(0, 1)`` (oracle :238-274)."""
    geometry = Geometry(kernel_size, stride)
    coarse = np.asarray(_as_coords(coarse_coordinates), dtype=np.int64).reshape(-1, 3)
    coarse_min = coarse.min(axis=0)
    shape = tuple(int(v) for v in (coarse.max(axis=0) - coarse_min + 1))
    dense = torch.zeros((1, features.shape[1], *shape), dtype=features.dtype)
    local = coarse - coarse_min
    dense[0, :, local[:, 0], local[:, 1], local[:, 2]] = features.t()
    values = torch.nn.functional.conv_transpose3d(dense, weights.transpose(0, 1).contiguous(), stride=geometry.stride, padding=0)
    origin = np.asarray(geometry.stride) * coarse_min - np.asarray(geometry.padding_before)
    return values, tuple(int(v) for v in origin)
