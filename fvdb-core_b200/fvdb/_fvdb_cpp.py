"""Drop-in for the convolution section of the reference's pybind11 module ``fvdb._fvdb_cpp``.

Same names, argument meaning and error behaviour as src/python/Bindings.cpp:491-674 and
src/python/GridBatchOps.cpp:764-792, implemented over the C ABI of libfvdbconv.so (include/fvdbconv.h).
PyTorch is used for device memory and streams only; every computation is a hand-written sm_100a kernel
reached through ``ctypes``.  There is no CPU path: CPU tensors are rejected.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import torch

from . import _lib
from ._lib import lib, check, i3

_DTYPE_CODE = {torch.float16: _lib.FVC_F16, torch.bfloat16: _lib.FVC_BF16, torch.float32: _lib.FVC_F32, torch.float64: _lib.FVC_F64}
_INT32_MAX = 2**31 - 1

# Executor variant forced by tests / bench ("auto" | "simt" | "tc"); the default picks per dtype / channels.
_FORCED_PATH = {"auto": 0, "simt": 1, "tc": 2}
_path = 0


def set_conv_path(name: str) -> None:
    """Force the CUDA-core ("simt") or tcgen05 ("tc") kernel family, or "auto" (default)."""
    global _path
    _path = _FORCED_PATH[name]


# Narrow half-precision layers (16 / 32 channels): grad_features AND grad_weights from one gather of grad_output.  Measured
# (profiles/r02_time_fused.jsonl, second block): 13 - 38 % faster than the two separate kernels on four of five narrow shapes and
# a tie on the fifth, once its MMA-issuing thread entered through elect.sync; FVC_FUSED_BACKWARD=0 restores the separate kernels.
_fused_backward = os.environ.get("FVC_FUSED_BACKWARD", "1") != "0"


def set_fused_backward(enabled: bool) -> None:
    """Experiment switch: serve narrow half-precision layers' backward with ONE kernel (fvc_conv_backward_fused) or with the
    separate weight-gradient and dgrad kernels."""
    global _fused_backward
    _fused_backward = bool(enabled)


def set_kernel_variant(variant: int, wgrad: bool = False) -> None:
    """Benchmark knob (fvc_set_tuning): pipeline-shape variant of the tensor-core forward (key 0) / weight-gradient (key 1)
    kernel; 0 = default."""
    check(lib.fvc_set_tuning(1 if wgrad else 0, int(variant)))


def _stream(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t: "torch.Tensor | None") -> int:
    return 0 if t is None or t.numel() == 0 else t.data_ptr()


def _require_cuda(device: torch.device, what: str) -> None:
    if device.type != "cuda":
        raise RuntimeError(
            f"{what}: this build executes the ConvolutionPlan path on CUDA (sm_100a) only and has no CPU fallback; got device {device}"
        )


def _vec3(value) -> list[int]:
    if isinstance(value, torch.Tensor):
        value = value.tolist()
    value = [int(v) for v in value]
    if len(value) != 3:
        raise ValueError(f"expected three values, got {value}")
    return value


# ---------------------------------------------------------------------------------------------
# ConvolutionGeometry (Bindings.cpp:497-540 over ConvolutionGeometry.h)
# ---------------------------------------------------------------------------------------------


class ConvolutionGeometry:
    """Immutable canonical geometry ``fine = stride * coarse + tap - padding_before``."""

    def __init__(self, kernel_size, stride):
        ks, st = _vec3(kernel_size), _vec3(stride)
        before, after, volume = i3([0, 0, 0]), i3([0, 0, 0]), C.c_int64(0)
        check(lib.fvc_geometry(i3(ks), i3(st), before, after, C.byref(volume)))
        self._kernel_size, self._stride = ks, st
        self._padding_before, self._padding_after = list(before), list(after)
        self._kernel_volume = int(volume.value)

    kernel_size = property(lambda self: list(self._kernel_size))
    stride = property(lambda self: list(self._stride))
    dilation = property(lambda self: [1, 1, 1])
    padding_before = property(lambda self: list(self._padding_before))
    padding_after = property(lambda self: list(self._padding_after))
    registration_offset = property(lambda self: [0, 0, 0])
    semantics_version = property(lambda self: 1)
    kernel_volume = property(lambda self: self._kernel_volume)
    phase_policy = property(lambda self: "torch_same_phase")

    def tap_coord(self, tap_index: int) -> list[int]:
        out = i3([0, 0, 0])
        check(lib.fvc_geometry_tap_coord(i3(self._kernel_size), int(tap_index), out))
        return list(out)

    def fine_from_coarse(self, coarse, tap) -> list[int]:
        out = i3([0, 0, 0])
        check(lib.fvc_geometry_fine_from_coarse(i3(self._kernel_size), i3(self._stride), i3(coarse), i3(tap), out))
        return list(out)

    def coarse_from_fine(self, fine, tap) -> "list[int] | None":
        out, ok = i3([0, 0, 0]), C.c_int32(0)
        check(lib.fvc_geometry_coarse_from_fine(i3(self._kernel_size), i3(self._stride), i3(fine), i3(tap), out, C.byref(ok)))
        return list(out) if ok.value else None


# ---------------------------------------------------------------------------------------------
# GridBatchData: the batched index grid (src/fvdb/GridBatchData.h:29-55) as torch-owned device arrays
# ---------------------------------------------------------------------------------------------


class GridBatchData:
    """Device index grid + per-grid metadata.  Identity (``is_same``) is object identity, as in the
    reference (src/python/GridBatchDataBinding.cpp:62-67)."""

    def __init__(self, *, device, num_grids, counts, leaves, lower, upper, root_keys, root_offsets, voxel_offsets, leaf_offsets, ijk, jidx, voxel_sizes, origins):
        self.device = torch.device(device)
        self.num_grids = int(num_grids)
        self.total_voxels, self.num_leaves, self.num_lower, self.num_upper = (int(c) for c in counts)
        self.leaves, self.lower, self.upper, self.root_keys = leaves, lower, upper, root_keys
        self.root_offsets, self.voxel_offsets, self.leaf_offsets = root_offsets, voxel_offsets, leaf_offsets
        self.ijk, self.jidx = ijk, jidx
        self.voxel_sizes = voxel_sizes  # float64 [B, 3], host (authoritative, GridBatchData.h:29-55)
        self.origins = origins  # float64 [B, 3], host
        self._struct = _lib.FvcGridBatch(
            self.num_grids, self.num_leaves, self.num_lower, self.num_upper, self.total_voxels,
            _ptr(leaves), _ptr(lower), _ptr(upper), _ptr(root_keys), root_offsets.data_ptr(), voxel_offsets.data_ptr(), leaf_offsets.data_ptr(),
        )

    @property
    def struct(self):
        return C.byref(self._struct)

    @property
    def grid_count(self) -> int:
        return self.num_grids

    def is_same(self, other: "GridBatchData") -> bool:
        return self is other


def _as_i32_coords(ijk: torch.Tensor) -> torch.Tensor:
    if ijk.dtype.is_floating_point or ijk.dtype == torch.bool:
        raise TypeError("ijk must have an integer type")
    if ijk.ndim != 2 or ijk.shape[1] != 3:
        raise ValueError(f"ijk must have shape (n, 3), got {tuple(ijk.shape)}")
    return ijk.to(torch.int32).contiguous()


def build_grid_from_ijk(ijk: torch.Tensor, jidx: "torch.Tensor | None", num_grids: int, voxel_sizes: torch.Tensor, origins: torch.Tensor) -> GridBatchData:
    """(grid, ijk) rows -> index grid (replaces ops/BuildGridFromIjk.cu:52-111).  Duplicates and negative
    coordinates are fine; one batched device pass + one count read-back."""
    device = ijk.device
    _require_cuda(device, "GridBatch construction")
    if num_grids > 1024:
        raise RuntimeError(f"batch size {num_grids} exceeds the 1024-grid limit")
    ijk = _as_i32_coords(ijk)
    n = int(ijk.shape[0])
    bidx = None
    if jidx is not None and num_grids > 1:
        bidx = jidx.to(device=device, dtype=torch.int32).contiguous()
    with torch.cuda.device(device):
        stream = _stream(device)
        scratch_bytes = int(lib.fvc_grid_build_scratch_bytes(n))
        scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=device)
        counts = (C.c_int64 * 4)()
        check(lib.fvc_grid_build_count(_ptr(ijk), _ptr(bidx), n, num_grids, scratch.data_ptr(), scratch_bytes, C.byref(counts), stream))
        nv, nl, nlow, nup = (int(c) for c in counts)
        leaves = torch.empty(nl * 128, dtype=torch.uint8, device=device)
        lower = torch.empty((nlow, 4096), dtype=torch.int32, device=device)
        upper = torch.empty((nup, 32768), dtype=torch.int32, device=device)
        root_keys = torch.empty((nup, 4), dtype=torch.int32, device=device)
        root_offsets = torch.empty(num_grids + 1, dtype=torch.int32, device=device)
        voxel_offsets = torch.empty(num_grids + 1, dtype=torch.int64, device=device)
        leaf_offsets = torch.empty(num_grids + 1, dtype=torch.int32, device=device)
        out_ijk = torch.empty((nv, 3), dtype=torch.int32, device=device)
        out_jidx = torch.empty(nv, dtype=torch.int32, device=device)
        check(
            lib.fvc_grid_build_fill(
                _ptr(ijk), _ptr(bidx), n, num_grids, scratch.data_ptr(), scratch_bytes, C.byref(counts),
                _ptr(leaves), _ptr(lower), _ptr(upper), _ptr(root_keys), root_offsets.data_ptr(), voxel_offsets.data_ptr(),
                leaf_offsets.data_ptr(), _ptr(out_ijk), _ptr(out_jidx), stream,
            )
        )
    return GridBatchData(
        device=device, num_grids=num_grids, counts=(nv, nl, nlow, nup), leaves=leaves, lower=lower, upper=upper, root_keys=root_keys,
        root_offsets=root_offsets, voxel_offsets=voxel_offsets, leaf_offsets=leaf_offsets, ijk=out_ijk, jidx=out_jidx,
        voxel_sizes=voxel_sizes, origins=origins,
    )


_LEAF_NEIGHBOURS = torch.tensor([[dx, dy, dz] for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)], dtype=torch.int32) * 8
use_leaf_morphology = True  # tests switch this off to cross-check the fast path against the candidate path


def _dilated_grid(grid: GridBatchData, lo: list[int], hi: list[int], voxel_sizes: torch.Tensor, origins: torch.Tensor) -> GridBatchData:
    """out(c) = OR over o in [lo, hi]^3 of grid(c - o), by leaf-mask morphology (csrc/grid_morph.cu): the neighbour-leaf origins
    (27 keys per LEAF) go through the ordinary builder, one warp per output leaf dilates the 3x3x3 source masks, a scan of the
    leaf counts gives the row bases.  One extra count read-back."""
    device = grid.device
    words = grid.leaves.view(torch.int32).reshape(-1, 32)  # FvcLeaf: mask 16 | prefix 4 | base | batch | origin[3] | count | ...
    origin, batch = words[:, 22:25], words[:, 21]
    near = [o for o in _LEAF_NEIGHBOURS.tolist() if all((d == 0) or (d < 0 and l < 0) or (d > 0 and h > 0) for d, l, h in zip(o, lo, hi))]
    near_t = torch.tensor(near, dtype=torch.int32, device=device)
    cand = (origin[:, None, :] + near_t[None]).reshape(-1, 3)
    skeleton = build_grid_from_ijk(cand, batch.repeat_interleave(len(near)), grid.num_grids, voxel_sizes, origins)
    nl = skeleton.num_leaves
    with torch.cuda.device(device):
        stream = _stream(device)
        counts = torch.empty(nl, dtype=torch.int32, device=device)
        check(lib.fvc_grid_dilate_leaves(grid.struct, _ptr(skeleton.leaves), nl, i3(lo), i3(hi), _ptr(counts), stream))
        ends = torch.cumsum(counts, 0, dtype=torch.int64)
        total = int(ends[-1]) if nl else 0  # the read-back that sizes the voxel list
        if total > _INT32_MAX:
            raise RuntimeError(f"generated topology would hold {total} voxels, exceeding the int32 limit")
        base = (ends - counts).to(torch.int32)
        out_ijk = torch.empty((total, 3), dtype=torch.int32, device=device)
        out_jidx = torch.empty(total, dtype=torch.int32, device=device)
        check(lib.fvc_grid_expand_leaves(_ptr(skeleton.leaves), nl, _ptr(base), _ptr(out_ijk), _ptr(out_jidx), stream))
        ends0 = torch.cat([ends.new_zeros(1), ends])
        voxel_offsets = ends0[skeleton.leaf_offsets.long()].contiguous()
    return GridBatchData(
        device=device, num_grids=grid.num_grids, counts=(total, nl, skeleton.num_lower, skeleton.num_upper), leaves=skeleton.leaves,
        lower=skeleton.lower, upper=skeleton.upper, root_keys=skeleton.root_keys, root_offsets=skeleton.root_offsets, voxel_offsets=voxel_offsets,
        leaf_offsets=skeleton.leaf_offsets, ijk=out_ijk, jidx=out_jidx, voxel_sizes=voxel_sizes, origins=origins,
    )


def _generated_grid(grid: GridBatchData, kernel_size, stride, transposed: bool) -> GridBatchData:
    ks, st = _vec3(kernel_size), _vec3(stride)
    geometry = ConvolutionGeometry(ks, st)  # validation (ValueError on non-positive sizes)
    device = grid.device
    _require_cuda(device, "conv_grid")
    scale = torch.tensor(st, dtype=torch.float64)
    voxel_sizes = grid.voxel_sizes / scale if transposed else grid.voxel_sizes * scale  # BuildGridForConv.cu:533-538, Transpose :350-355
    if use_leaf_morphology and st == [1, 1, 1] and max(ks) <= 8 and grid.total_voxels > 0:
        # stride 1: the support is a box dilation.  forward: coarse = fine - tap + pad; transposed: fine = coarse + tap - pad
        pad = geometry.padding_before
        lo = [-p for p in pad] if transposed else [p - (k - 1) for p, k in zip(pad, ks)]
        hi = [k - 1 - p for p, k in zip(pad, ks)] if transposed else list(pad)
        return _dilated_grid(grid, lo, hi, voxel_sizes, grid.origins.clone())
    with torch.cuda.device(device):
        stream = _stream(device)
        n = grid.total_voxels
        counter = torch.zeros(2, dtype=torch.int64, device=device)
        count = C.c_int64(0)
        check(lib.fvc_conv_grid_count(_ptr(grid.ijk), n, i3(ks), i3(st), int(transposed), counter.data_ptr(), C.byref(count), stream))
        m = int(count.value)
        if m > _INT32_MAX:
            raise RuntimeError(
                f"generated topology would stage {m} candidate coordinates ({n} input voxels * {ks[0] * ks[1] * ks[2]} kernel taps), "
                "exceeding the int32 limit; provide an explicit target grid"
            )
        cand_ijk = torch.empty((m, 3), dtype=torch.int32, device=device)
        cand_bidx = torch.empty(m, dtype=torch.int32, device=device)
        check(
            lib.fvc_conv_grid_emit(
                _ptr(grid.ijk), _ptr(grid.jidx), n, i3(ks), i3(st), int(transposed), m, _ptr(cand_ijk), _ptr(cand_bidx), counter.data_ptr() + 8, stream
            )
        )
    return build_grid_from_ijk(cand_ijk, cand_bidx, grid.num_grids, voxel_sizes, grid.origins.clone())


def conv_grid(grid: GridBatchData, kernel_size: Sequence[int], stride: Sequence[int]) -> GridBatchData:
    """Complete forward support (GridBatchOps.cpp:764-776 -> ops::buildGridForConv)."""
    return _generated_grid(grid, kernel_size, stride, transposed=False)


def conv_transpose_grid(grid: GridBatchData, kernel_size: Sequence[int], stride: Sequence[int]) -> GridBatchData:
    """Complete transposed support (GridBatchOps.cpp:778-792 -> ops::buildGridForConvTranspose)."""
    return _generated_grid(grid, kernel_size, stride, transposed=True)


def neighbor_indexes(grid: GridBatchData, ijk: torch.Tensor, jidx: "torch.Tensor | None", extent: int, bitshift: int = 0) -> torch.Tensor:
    _require_cuda(grid.device, "neighbor_indexes")
    ijk = _as_i32_coords(ijk.to(grid.device))
    nq, w = int(ijk.shape[0]), 2 * int(extent) + 1
    bidx = None if jidx is None or grid.num_grids == 1 else jidx.to(device=grid.device, dtype=torch.int32).contiguous()
    out = torch.empty((nq, w, w, w), dtype=torch.int64, device=grid.device)
    with torch.cuda.device(grid.device):
        check(lib.fvc_neighbor_indexes(grid.struct, _ptr(ijk), _ptr(bidx), nq, int(extent), int(bitshift), _ptr(out), _stream(grid.device)))
    return out


def ijk_to_index(grid: GridBatchData, ijk: torch.Tensor, jidx: "torch.Tensor | None", cumulative: bool = False) -> torch.Tensor:
    _require_cuda(grid.device, "ijk_to_index")
    ijk = _as_i32_coords(ijk.to(grid.device))
    nq = int(ijk.shape[0])
    bidx = None if jidx is None or grid.num_grids == 1 else jidx.to(device=grid.device, dtype=torch.int32).contiguous()
    out = torch.empty(nq, dtype=torch.int64, device=grid.device)
    with torch.cuda.device(grid.device):
        check(lib.fvc_ijk_to_index(grid.struct, _ptr(ijk), _ptr(bidx), nq, int(cumulative), _ptr(out), _stream(grid.device)))
    return out


# ---------------------------------------------------------------------------------------------
# GatherScatterDefaultTopology (Bindings.cpp:542-561 over GatherScatterDefault.h:59-81)
# ---------------------------------------------------------------------------------------------


def _map_pitch(rows: int) -> int:
    """Row pitch of a tap-major dense map: whole 128-row tiles, so the executors can stream 16-byte map chunks."""
    return max((rows + 127) // 128 * 128, 128)


def _tile_mask(nbr: torch.Tensor, rows: int, kernel_volume: int) -> "torch.Tensor | None":
    """Per 128-row tile tap bitmask of a dense map (fvc_kmap_tile_mask); lets the executors skip empty units."""
    if rows == 0 or kernel_volume == 0 or kernel_volume > 4096:
        return None
    words = (kernel_volume + 63) // 64
    mask = torch.empty(((rows + 127) // 128) * words, dtype=torch.int64, device=nbr.device)
    with torch.cuda.device(nbr.device):
        check(lib.fvc_kmap_tile_mask(_ptr(nbr), int(nbr.shape[1]), rows, kernel_volume, mask.data_ptr(), _stream(nbr.device)))
    return mask


class _MapCore:
    """Storage shared by a topology and its constant-time reversed view.

    ``nbr`` is the output-stationary tap-major dense map of the *built* direction ([K^3, pitch] int32, feature row or -1) and
    ``mask`` its per-tile tap bitmask, both written by ONE kernel with no host synchronisation.  Everything else is derived
    lazily, on first use: ``nbr_rev`` (input-stationary; serves dgrad of the built direction and the forward pass of the
    reversed view), the host offsets (the one D2H read of the tap counts, reference: GatherScatterDefault.cu:149) and the
    reference's CSR-by-tap arrays -- the tensor-core executors read only the dense maps, so a training step never builds them.
    """

    def __init__(self, nbr, mask, tap_counts, n_feature, n_output, kernel_volume, symmetric=False):
        # symmetric: same grid on both sides, stride 1, odd kernel -> nbr_rev[k] == nbr[K^3 - 1 - k] (row i reaches o through tap k
        # iff o reaches i through the mirrored tap), so dgrad can run on `nbr` itself with the taps of W mirrored
        self.symmetric = bool(symmetric)
        self.nbr, self.mask, self._tap_counts = nbr, mask, tap_counts
        self._nbr_rev, self._mask_rev = None, None
        self._offsets_host, self._offsets_dev, self._gather, self._scatter = None, None, None, None
        self.n_feature, self.n_output, self.kernel_volume = n_feature, n_output, kernel_volume

    # ---- lazily derived views ---------------------------------------------------------------------
    @property
    def offsets_host(self) -> torch.Tensor:
        if self._offsets_host is None:
            offsets = torch.zeros(self.kernel_volume + 1, dtype=torch.int64)
            offsets[1:] = torch.cumsum(self._tap_counts.cpu(), 0)  # the one D2H read (synchronises), only when somebody asks
            self._offsets_host = offsets
        return self._offsets_host

    @property
    def total_pairs(self) -> int:
        return int(self.offsets_host[-1]) if self.kernel_volume else 0

    def _csr(self) -> None:
        if self._gather is not None:
            return
        device, k3, n_out = self.nbr.device, self.kernel_volume, self.n_output
        total = self.total_pairs
        with torch.cuda.device(device):
            stream = _stream(device)
            gather = torch.empty(total, dtype=torch.int32, device=device)
            scatter = torch.empty(total, dtype=torch.int32, device=device)
            offsets_dev = torch.empty(k3 + 1, dtype=torch.int64, device=device)
            scratch_bytes = int(lib.fvc_kmap_csr_scratch_bytes(n_out, k3))
            scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=device)
            check(
                lib.fvc_kmap_to_csr(
                    _ptr(self.nbr), int(self.nbr.shape[1]), n_out, k3, self._tap_counts.data_ptr(), offsets_dev.data_ptr(), _ptr(gather), _ptr(scatter),
                    scratch.data_ptr(), scratch_bytes, stream,
                )
            )
        self._gather, self._scatter, self._offsets_dev = gather, scatter, offsets_dev

    @property
    def gather(self) -> torch.Tensor:
        self._csr()
        return self._gather

    @property
    def scatter(self) -> torch.Tensor:
        self._csr()
        return self._scatter

    @property
    def offsets_dev(self) -> torch.Tensor:
        self._csr()
        return self._offsets_dev

    def nbr_rev(self) -> torch.Tensor:
        if self._nbr_rev is None:
            device = self.nbr.device
            pitch = _map_pitch(self.n_feature)
            rev = torch.empty((self.kernel_volume, pitch), dtype=torch.int32, device=device)
            with torch.cuda.device(device):
                check(
                    lib.fvc_kmap_reverse_from_dense(
                        _ptr(self.nbr), int(self.nbr.shape[1]), self.n_output, self.kernel_volume, self.n_feature, _ptr(rev), pitch, _stream(device)
                    )
                )
            self._nbr_rev = rev
            self._mask_rev = _tile_mask(rev, self.n_feature, self.kernel_volume)
        return self._nbr_rev

    def mask_rev(self) -> "torch.Tensor | None":
        self.nbr_rev()
        return self._mask_rev

    def zero_degree_outputs(self) -> int:
        """Output rows no tap reaches (coverage policy, reference convolution_plan.py:341-345), from the dense map."""
        if self.n_output == 0:
            return 0
        degree = torch.empty(self.n_output, dtype=torch.int32, device=self.nbr.device)
        with torch.cuda.device(self.nbr.device):
            check(lib.fvc_kmap_degree(_ptr(self.nbr), int(self.nbr.shape[1]), self.n_output, self.kernel_volume, degree.data_ptr(), _stream(self.nbr.device)))
        return int((degree == 0).sum())


class GatherScatterDefaultTopology:
    """CSR-by-tap kernel map; same read-only properties as the reference binding."""

    def __init__(self, core: _MapCore, kernel_size, stride, is_transposed: bool, reversed_view: bool):
        self._core, self._reversed = core, reversed_view
        self._kernel_size, self._stride, self._is_transposed = list(kernel_size), list(stride), bool(is_transposed)

    gather_indices = property(lambda self: self._core.scatter if self._reversed else self._core.gather)
    scatter_indices = property(lambda self: self._core.gather if self._reversed else self._core.scatter)
    offsets = property(lambda self: self._core.offsets_host)  # int64 [K^3 + 1] on the HOST, as in the reference (GatherScatterDefault.h:70)
    feature_total_voxels = property(lambda self: self._core.n_output if self._reversed else self._core.n_feature)
    output_total_voxels = property(lambda self: self._core.n_feature if self._reversed else self._core.n_output)
    kernel_volume = property(lambda self: self._core.kernel_volume)
    total_pairs = property(lambda self: self._core.total_pairs)
    kernel_size = property(lambda self: list(self._kernel_size))
    stride = property(lambda self: list(self._stride))
    is_transposed = property(lambda self: self._is_transposed)

    # engine-private views
    def _out_map(self) -> torch.Tensor:  # [K^3, pitch]: for each output row and tap, the feature row
        return self._core.nbr_rev() if self._reversed else self._core.nbr

    def _in_map(self) -> torch.Tensor:  # [K^3, pitch]: for each feature row and tap, the output row
        return self._core.nbr if self._reversed else self._core.nbr_rev()

    def _out_mask(self) -> "torch.Tensor | None":
        return self._core.mask_rev() if self._reversed else self._core.mask

    def _in_mask(self) -> "torch.Tensor | None":
        return self._core.mask if self._reversed else self._core.mask_rev()

    def _dgrad_plan(self) -> "tuple[torch.Tensor, torch.Tensor | None, bool]":
        """(input-stationary map, its tile mask, mirror the taps of W?) for dgrad.  A symmetric map serves dgrad as it is --
        no reversed copy is ever built (half the map memory, a shorter plan build)."""
        if self._core.symmetric and not self._reversed:
            return self._core.nbr, self._core.mask, True
        return self._in_map(), self._in_mask(), False

    @property
    def device(self) -> torch.device:
        return self._core.nbr.device


def _build_topology(feature_grid: GridBatchData, output_grid: GridBatchData, kernel_size, stride, transposed: bool) -> GatherScatterDefaultTopology:
    ks, st = _vec3(kernel_size), _vec3(stride)
    if feature_grid.device != output_grid.device:  # GatherScatterDefault.cu:58-62
        raise RuntimeError(f"feature_grid and output_grid must be on the same device, got {feature_grid.device} and {output_grid.device}")
    device = output_grid.device
    _require_cuda(device, "gs_build_topology")
    k3 = ks[0] * ks[1] * ks[2]
    n_out, n_feat = output_grid.total_voxels, feature_grid.total_voxels
    pitch = _map_pitch(n_out)
    with torch.cuda.device(device):
        stream = _stream(device)
        nbr = torch.empty((k3, pitch), dtype=torch.int32, device=device)
        tap_counts = torch.empty(k3, dtype=torch.int64, device=device)
        mask = None
        if n_out > 0 and 0 < k3 <= 4096:  # the tile tap-mask comes out of the same kernel
            mask = torch.empty(((n_out + 127) // 128) * ((k3 + 63) // 64), dtype=torch.int64, device=device)
        # ONE kernel, no host synchronisation: dense map + per-tap pair counts + tile tap-masks
        check(lib.fvc_kmap_build(feature_grid.struct, output_grid.struct, i3(ks), i3(st), int(transposed), _ptr(nbr), pitch, tap_counts.data_ptr(), _ptr(mask), stream))
    symmetric = feature_grid.is_same(output_grid) and st == [1, 1, 1] and all(k % 2 == 1 for k in ks)
    core = _MapCore(nbr, mask, tap_counts, n_feat, n_out, k3, symmetric=symmetric)
    return GatherScatterDefaultTopology(core, ks, st, transposed, reversed_view=False)


def gs_build_topology(feature_grid, output_grid, kernel_size, stride) -> GatherScatterDefaultTopology:
    """Forward kernel map (Bindings.cpp:570-583 -> gatherScatterDefaultSparseConvTopology)."""
    return _build_topology(feature_grid, output_grid, kernel_size, stride, transposed=False)


def gs_build_transpose_topology(feature_grid, output_grid, kernel_size, stride) -> GatherScatterDefaultTopology:
    """Transposed kernel map (Bindings.cpp:612-625 -> gatherScatterDefaultSparseConvTransposeTopology)."""
    return _build_topology(feature_grid, output_grid, kernel_size, stride, transposed=True)


def gs_reverse_topology(topology: GatherScatterDefaultTopology) -> GatherScatterDefaultTopology:
    """Constant-time reversed view aliasing the same tensors (GatherScatterDefault.cu:273-294)."""
    return GatherScatterDefaultTopology(topology._core, topology._kernel_size, topology._stride, not topology._is_transposed, not topology._reversed)


def validate_gather_scatter_default_topology(fine_grid: GridBatchData, coarse_grid: GridBatchData, topology, *, gather_indices=None, scatter_indices=None,
                                             offsets=None, feature_total_voxels=None, output_total_voxels=None) -> None:
    """Explicit test / debug validation of a kernel map (validateGatherScatterDefaultTopology, GatherScatterDefault.cu:342-527):
    metadata, offsets, index ranges, every stored edge against the canonical relation ``fine = S * coarse + tap - pad``
    inside one batch item, no duplicate edge, and stored edge set == complete canonical relation.  Raises RuntimeError with
    the reference's messages.  The keyword overrides stand in for the corrupted copies the reference test builds
    (src/tests/GatherScatterDefaultConvTest.cu:876-930).  Runs as torch ops + the native lookup kernel on the grids' device."""
    if fine_grid.device != coarse_grid.device:
        raise RuntimeError(f"fine_grid and coarse_grid must be on the same device, got {fine_grid.device} and {coarse_grid.device}")
    if fine_grid.num_grids != coarse_grid.num_grids:
        raise RuntimeError(f"fine_grid and coarse_grid batch sizes must match, got {fine_grid.num_grids} and {coarse_grid.num_grids}")
    forward = not topology.is_transposed
    feature_grid, output_grid = (fine_grid, coarse_grid) if forward else (coarse_grid, fine_grid)
    n_feat = topology.feature_total_voxels if feature_total_voxels is None else feature_total_voxels
    n_out = topology.output_total_voxels if output_total_voxels is None else output_total_voxels
    if n_feat != feature_grid.total_voxels:
        raise RuntimeError(f"topology feature voxel count {n_feat} does not match its {'fine' if forward else 'coarse'} domain count {feature_grid.total_voxels}")
    if n_out != output_grid.total_voxels:
        raise RuntimeError(f"topology output voxel count {n_out} does not match its {'coarse' if forward else 'fine'} domain count {output_grid.total_voxels}")
    geometry = ConvolutionGeometry(topology.kernel_size, topology.stride)
    k3 = topology.kernel_volume
    if k3 != geometry.kernel_volume:
        raise RuntimeError(f"topology kernel volume {k3} does not match geometry volume {geometry.kernel_volume}")
    gather = topology.gather_indices if gather_indices is None else gather_indices
    scatter = topology.scatter_indices if scatter_indices is None else scatter_indices
    offs = topology.offsets if offsets is None else offsets
    total = topology.total_pairs
    for tensor, name in ((gather, "gather_indices"), (scatter, "scatter_indices")):
        if tensor.dim() != 1:
            raise RuntimeError(f"{name} must be one-dimensional")
        if tensor.dtype != torch.int32:
            raise RuntimeError(f"{name} must have int32 dtype")
        if tensor.device != fine_grid.device:
            raise RuntimeError(f"{name} must be on grid device {fine_grid.device}, got {tensor.device}")
        if not tensor.is_contiguous():
            raise RuntimeError(f"{name} must be contiguous")
        if tensor.numel() != total:
            raise RuntimeError(f"{name} length {tensor.numel()} does not match total pair count {total}")
    if offs.dim() != 1 or offs.dtype != torch.int64 or offs.device.type != "cpu" or not offs.is_contiguous():
        raise RuntimeError("offsets must be a contiguous one-dimensional int64 tensor stored on CPU")
    if offs.numel() != k3 + 1:
        raise RuntimeError(f"offset length {offs.numel()} does not match kernel volume + 1 ({k3 + 1})")
    o = offs.tolist()
    if o[0] != 0:
        raise RuntimeError(f"offsets must start at zero, got {o[0]}")
    for tap in range(k3):
        if o[tap] > o[tap + 1]:
            raise RuntimeError(f"offsets must be monotone at tap {tap}: {o[tap]} > {o[tap + 1]}")
    if o[k3] != total:
        raise RuntimeError(f"final offset {o[k3]} does not match total pair count {total}")
    g64, s64 = gather.long(), scatter.long()
    for idx, name, limit in ((g64, "gather", n_feat), (s64, "scatter", n_out)):
        bad = ((idx < 0) | (idx >= limit)).nonzero()
        if bad.numel():
            pair = int(bad[0])
            raise RuntimeError(f"{name} index out of range at pair {pair}: {int(idx[pair])}")
    device = fine_grid.device
    fine_idx, coarse_idx = (g64, s64) if forward else (s64, g64)
    tap_of_pair = torch.repeat_interleave(torch.arange(k3, device=device), torch.tensor([o[t + 1] - o[t] for t in range(k3)], device=device)) if total else g64
    ks, st, pad = geometry.kernel_size, geometry.stride, geometry.padding_before
    taps = torch.tensor([geometry.tap_coord(t) for t in range(k3)], dtype=torch.int32, device=device).reshape(k3, 3)
    scale, shift = torch.tensor(st, dtype=torch.int32, device=device), torch.tensor(pad, dtype=torch.int32, device=device)
    if total:
        crossing = (fine_grid.jidx.long()[fine_idx] != coarse_grid.jidx.long()[coarse_idx]).nonzero()
        if crossing.numel():
            raise RuntimeError(f"edge crosses batch domains at pair {int(crossing[0])}")
        want_fine = coarse_grid.ijk[coarse_idx] * scale + taps[tap_of_pair] - shift  # fineFromCoarse (ConvolutionGeometry.h:99-104)
        wrong = (fine_grid.ijk[fine_idx] != want_fine).any(dim=1).nonzero()
        if wrong.numel():
            pair = int(wrong[0])
            raise RuntimeError(f"edge does not satisfy canonical fine/coarse geometry at pair {pair}, tap {int(tap_of_pair[pair])}")
        key = (tap_of_pair * max(n_feat, 1) + (g64 if forward else s64)) * max(n_out, 1) + (s64 if forward else g64)
        uniq, counts = torch.unique(key, return_counts=True)
        if uniq.numel() != total:
            raise RuntimeError("duplicate fine/coarse/tap edge in the stored topology")
    # the complete canonical relation: every coarse voxel probes every tap on the fine grid (same batch item)
    expected = 0
    nc = coarse_grid.total_voxels
    if nc and k3 and fine_grid.total_voxels:
        for tap in range(k3):
            probe = coarse_grid.ijk * scale + taps[tap] - shift
            hit = ijk_to_index(fine_grid, probe, coarse_grid.jidx, cumulative=True)
            expected += int((hit >= 0).sum())
    if expected != total:
        raise RuntimeError(f"stored topology edge set does not equal the complete canonical relation: got {total} edges, expected {expected}")


# ---------------------------------------------------------------------------------------------
# Execution (Bindings.cpp:585-653 over GatherScatterDefault.cu:635-924)
# ---------------------------------------------------------------------------------------------


def _check_conv(features: torch.Tensor, weights: torch.Tensor, topo: GatherScatterDefaultTopology, name: str) -> None:
    # GatherScatterDefault.cu:635-667
    if features.dim() != 2:
        raise RuntimeError(f"{name}: features must be 2D")
    if features.size(0) != topo.feature_total_voxels:
        raise RuntimeError(f"{name}: features.size(0)={features.size(0)} must match featureTotalVoxels={topo.feature_total_voxels}")
    if not features.is_floating_point():
        raise RuntimeError(f"{name}: features must be floating point")
    if not features.is_contiguous():
        raise RuntimeError(f"{name}: features must be contiguous")
    if weights.dim() != 5:
        raise RuntimeError(f"{name}: weights must be 5D [C_out, C_in, k0, k1, k2]")
    if not weights.is_floating_point():
        raise RuntimeError(f"{name}: weights must be floating point")
    if features.size(1) != weights.size(1):
        raise RuntimeError(f"{name}: features channels={features.size(1)} must match weights C_in={weights.size(1)}")
    if list(weights.shape[2:]) != topo.kernel_size:
        raise RuntimeError(f"{name}: weights spatial dims must match topology kernel_size")
    if features.device != weights.device:
        raise RuntimeError(f"{name}: features and weights must be on the same device")
    _require_cuda(features.device, name)
    if features.device != topo.device:
        raise RuntimeError(f"{name}: features and topology must be on the same device")


def _working_dtype(features: torch.Tensor, weights: torch.Tensor) -> torch.dtype:
    working = torch.result_type(features, weights)  # promoteFloatTypes, GatherScatterDefault.cu:592-595
    if working not in _DTYPE_CODE:
        raise RuntimeError(f"no convolution kernel registered for dtype {working}")
    return working


def _pack_weights(weights: torch.Tensor, working: torch.dtype, layout: int, flip_taps: bool = False) -> torch.Tensor:
    """Public ``[Cout,Cin,k0,k1,k2]`` (any strides) -> ``[K,Cin,Cout]`` (layout 0) / ``[K,Cout,Cin]`` (layout 1) in ``working``."""
    cout, cin, k0, k1, k2 = weights.shape
    out = torch.empty((k0 * k1 * k2, cin, cout) if layout == 0 else (k0 * k1 * k2, cout, cin), dtype=working, device=weights.device)
    strides = (C.c_int64 * 5)(*weights.stride())
    check(
        lib.fvc_pack_weights(
            _ptr(weights), C.byref(strides), _DTYPE_CODE[weights.dtype], cout, cin, k0, k1, k2, layout, int(flip_taps), _DTYPE_CODE[working], _ptr(out), _stream(weights.device)
        )
    )
    return out


def _prepare_weights(weights: torch.Tensor, working: torch.dtype, transpose: bool, flip_taps: bool = False) -> torch.Tensor:
    """The operand the executor consumes (fvc_conv_prepare_weights): ONE launch from the public layout to the tcgen05
    shared-memory image (or the packed array of the CUDA-core kernels).  ``transpose``: W[k]^T for dgrad."""
    cout, cin, k0, k1, k2 = (int(v) for v in weights.shape)
    k3, code = k0 * k1 * k2, _DTYPE_CODE[working]
    cin_e, cout_e = (cout, cin) if transpose else (cin, cout)
    nbytes = int(lib.fvc_conv_weights_bytes(cin_e, cout_e, k3, code, _path))
    blob = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=weights.device)
    strides = (C.c_int64 * 5)(*weights.stride())
    check(
        lib.fvc_conv_prepare_weights(
            _ptr(weights), C.byref(strides), _DTYPE_CODE[weights.dtype], cout, cin, k0, k1, k2, int(transpose), int(flip_taps), code, _path,
            blob.data_ptr(), nbytes, _stream(weights.device),
        )
    )
    return blob


def split_rows(x: torch.Tensor) -> torch.Tensor:
    """fp32 rows ``[n, c]`` -> bf16 split rows ``[n, 3, c]`` (fvc_split_rows): the operand of the fp32 tensor-core kernels.  A
    layer splits its input once and reuses the rows for forward and weight gradient."""
    n, c = x.shape
    out = torch.empty((n, 3, c), dtype=torch.bfloat16, device=x.device)
    if n:
        check(lib.fvc_split_rows(x.data_ptr(), n, c, out.data_ptr(), _stream(x.device)))
    return out


def _tensor_core_fp32(cin: int, cout: int, k3: int) -> bool:
    return int(lib.fvc_conv_kernel_family(cin, cout, k3, _lib.FVC_F32, _path, 0)) == 2


class ConvStats:
    """Per-block column sums of a convolution output written by the fused epilogue (FvcConvEpilogue.stats)."""

    def __init__(self, partial: torch.Tensor, rows_per_block: int, rows: int):
        self.partial, self.rows_per_block, self.rows = partial, rows_per_block, rows


def _run_conv(x: torch.Tensor, w_prepared: torch.Tensor, nbr: torch.Tensor, n_in: int, n_out: int, cin: int, cout: int, k3: int,
              bias: "torch.Tensor | None" = None, tile_mask: "torch.Tensor | None" = None, *, dtype: "torch.dtype | None" = None, x_is_split: bool = False,
              scale: "torch.Tensor | None" = None, shift: "torch.Tensor | None" = None, residual: "torch.Tensor | None" = None, relu: "bool | int" = False,
              want_stats: bool = False):
    """One output-stationary pass (forward, or dgrad on the reversed map) over prepared weights; optional fused epilogue
    ``act2(act1((acc + bias) * scale + shift) + residual)``: ``relu`` = True / 1 -> act1 (before the residual), 2 -> act2 (after), 3 -> both."""
    device = x.device
    dtype = dtype or x.dtype
    y = torch.empty((n_out, cout), dtype=dtype, device=device)
    code = _DTYPE_CODE[dtype]
    stats = None
    if want_stats:
        rpb = C.c_int32(0)
        blocks = int(lib.fvc_conv_stats_blocks(n_out, cin, cout, k3, code, _path, C.byref(rpb)))
        if rpb.value == 0:
            raise RuntimeError("fused convolution statistics need the tensor-core path")
        stats = ConvStats(torch.empty((max(blocks, 1), 2, cout), dtype=torch.float32, device=device), int(rpb.value), n_out)
    if n_out == 0:
        return (y, stats) if want_stats else y
    scratch_bytes = int(lib.fvc_conv_scratch_bytes(n_in, n_out, cin, cout, k3, code)) if (dtype == torch.float32 and not x_is_split) else 0
    scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=device) if scratch_bytes else None
    epi = _lib.FvcConvEpilogue(_ptr(bias) or None, _ptr(scale) or None, _ptr(shift) or None, _ptr(residual) or None, int(relu),
                               stats.partial.data_ptr() if stats is not None else None)
    check(
        lib.fvc_conv_forward_ex(
            _ptr(x), int(x_is_split), w_prepared.data_ptr(), C.byref(epi), y.data_ptr(), _ptr(nbr), int(nbr.shape[1]), _ptr(tile_mask), n_in, n_out, cin, cout, k3,
            code, _path, _ptr(scratch), scratch_bytes, _stream(device),
        )
    )
    return (y, stats) if want_stats else y


def _forward(features, weights, topo, name, want_transposed, bias=None, *, features_split=None, scale=None, shift=None, residual=None, relu=False,
             want_stats=False):
    _check_conv(features, weights, topo, name)
    if topo.is_transposed != want_transposed:
        raise RuntimeError(f"{name} requires topology with direction={'Transposed' if want_transposed else 'Forward'}")
    working = _working_dtype(features, weights)
    if features.dtype != working:
        features = features.to(working)  # :850-852
    cout, cin = int(weights.shape[0]), int(weights.shape[1])
    with torch.cuda.device(features.device):
        w = _prepare_weights(weights, working, transpose=False)
        if bias is not None:
            bias = bias.to(device=features.device, dtype=working).contiguous()
        if scale is not None:
            scale = scale.to(device=features.device, dtype=torch.float32).contiguous()
        if shift is not None:
            shift = shift.to(device=features.device, dtype=torch.float32).contiguous()
        if residual is not None:
            if residual.shape != (topo.output_total_voxels, cout):
                raise RuntimeError(f"{name}: residual must have shape {(topo.output_total_voxels, cout)}, got {tuple(residual.shape)}")
            residual = residual.to(working).contiguous()
        x, x_is_split = features, False
        if features_split is not None and working == torch.float32:
            x, x_is_split = features_split, True
        return _run_conv(x, w, topo._out_map(), topo.feature_total_voxels, topo.output_total_voxels, cin, cout, topo.kernel_volume, bias, topo._out_mask(),
                         dtype=working, x_is_split=x_is_split, scale=scale, shift=shift, residual=residual, relu=relu, want_stats=want_stats)


def _backward(grad_output, features, weights, topo, name, want_transposed, on_grad_weights=None, *, features_split=None, need_grad_features=True):
    _check_conv(features, weights, topo, name)
    if topo.is_transposed != want_transposed:
        raise RuntimeError(f"{name} requires {'direction=Transposed' if want_transposed else 'topology with direction=Forward'}")
    if grad_output.dim() != 2 or grad_output.size(0) != topo.output_total_voxels:
        raise RuntimeError("grad_output shape mismatch")  # :869-870
    if not grad_output.is_contiguous():
        raise RuntimeError("grad_output must be contiguous")
    if not grad_output.is_floating_point():
        raise RuntimeError("grad_output must be floating point")
    working = _working_dtype(features, weights)
    features = features.to(working) if features.dtype != working else features
    grad_output = grad_output.to(working) if grad_output.dtype != working else grad_output
    cout, cin = int(weights.shape[0]), int(weights.shape[1])
    k3, n_feat, n_out = topo.kernel_volume, topo.feature_total_voxels, topo.output_total_voxels
    if grad_output.size(1) != cout:
        raise RuntimeError("grad_output shape mismatch")
    device = features.device
    code = _DTYPE_CODE[working]
    with torch.cuda.device(device):
        # wgrad first: dW[k] = X[g]^T . dY[s]  (GatherScatterDefault.cu:806-813).  Its result is the only thing a data-parallel
        # job exchanges, so `on_grad_weights` (e.g. an asynchronous NCCL all-reduce) can overlap the dgrad kernel below.
        grad_weights = torch.empty(tuple(weights.shape), dtype=working, device=device)
        out_map = topo._out_map()
        # the tensor-core weight gradient reads the dense map only: neither the host offsets (a D2H sync) nor the CSR arrays are
        # ever materialised on that path
        tc_wgrad = k3 > 0 and int(lib.fvc_conv_kernel_family(cin, cout, k3, code, _path, 1)) == 2
        empty = n_feat == 0 or n_out == 0 or k3 == 0 or (not tc_wgrad and topo.total_pairs == 0)
        # fp32 on the tensor pipe: X (kept from the forward call when the caller has it) and dY are split ONCE each and the
        # split rows feed both the weight gradient and dgrad
        tc32 = working == torch.float32 and not empty and tc_wgrad and _tensor_core_fp32(cin, cout, k3) and _tensor_core_fp32(cout, cin, k3)
        x_op, x_split = (features_split, 1) if (tc32 and features_split is not None) else (features, 0)
        dy_op, dy_split = (split_rows(grad_output), 1) if tc32 else (grad_output, 0)
        # narrow half-precision layers: grad_features AND grad_weights from one gather of grad_output (conv_tc_bwd.cu)
        fused = (_fused_backward and need_grad_features and not empty and working in (torch.float16, torch.bfloat16)
                 and int(lib.fvc_conv_kernel_family(cin, cout, k3, code, _path, 2)) == 2)
        grad_features = None
        if empty:
            grad_weights.zero_()  # :771-777
        elif fused:
            in_map, in_mask, mirror = topo._dgrad_plan()
            wt = _prepare_weights(weights, working, transpose=True, flip_taps=mirror)
            grad_features = torch.empty((n_feat, cin), dtype=working, device=device)
            scratch_bytes = int(lib.fvc_conv_backward_fused_scratch_bytes(n_feat, cin, cout, k3))
            scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=device)
            check(
                lib.fvc_conv_backward_fused(
                    _ptr(grad_output), _ptr(features), wt.data_ptr(), _ptr(in_map), int(in_map.shape[1]), _ptr(in_mask), n_feat, n_out, cin, cout, k3, code,
                    int(mirror), grad_features.data_ptr(), grad_weights.data_ptr(), scratch.data_ptr(), scratch_bytes, _stream(device),
                )
            )
        elif tc_wgrad:
            scratch_bytes = int(lib.fvc_conv_wgrad_scratch_bytes(n_feat, n_out, 0, cin, cout, k3, code, _path, 1))
            scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=device) if scratch_bytes else None
            check(
                lib.fvc_conv_wgrad_ex(
                    _ptr(x_op), x_split, _ptr(dy_op), dy_split, None, None, None, None, _ptr(out_map), int(out_map.shape[1]),
                    _ptr(topo._out_mask()), n_feat, n_out, cin, cout, k3, code, _path, _ptr(grad_weights), _ptr(scratch), scratch_bytes, _stream(device),
                )
            )
        else:
            offsets_host = topo.offsets
            max_tap = int((offsets_host[1:] - offsets_host[:-1]).max())
            scratch_bytes = int(lib.fvc_conv_wgrad_scratch_bytes(n_feat, n_out, max_tap, cin, cout, k3, code, _path, 1))
            scratch = torch.empty(scratch_bytes, dtype=torch.uint8, device=device) if scratch_bytes else None
            check(
                lib.fvc_conv_wgrad_ex(
                    _ptr(x_op), x_split, _ptr(dy_op), dy_split, _ptr(topo.gather_indices), _ptr(topo.scatter_indices),
                    C.cast(offsets_host.data_ptr(), C.POINTER(C.c_int64)), topo._core.offsets_dev.data_ptr(), _ptr(out_map), int(out_map.shape[1]),
                    _ptr(topo._out_mask()), n_feat, n_out, cin, cout, k3, code, _path, _ptr(grad_weights), _ptr(scratch), scratch_bytes, _stream(device),
                )
            )
        if on_grad_weights is not None:
            on_grad_weights(grad_weights)
        # dgrad: dX[i] = sum_k dY[in_map[k][i]] . W[k]^T  (:803-804), output-stationary over features
        if not need_grad_features or fused:
            pass
        elif empty:
            grad_features = torch.zeros((n_feat, cin), dtype=working, device=device)  # :771-777
        else:
            in_map, in_mask, mirror = topo._dgrad_plan()
            wt = _prepare_weights(weights, working, transpose=True, flip_taps=mirror)
            grad_features = _run_conv(dy_op, wt, in_map, n_out, n_feat, cout, cin, k3, None, in_mask, dtype=working, x_is_split=bool(dy_split))
    return grad_features, grad_weights


def gs_conv(features, weights, topology, bias=None, **fused):
    """Forward sparse convolution (Bindings.cpp:585-594).  Extensions (keyword-only): ``features_split`` (fp32 split rows of
    ``features``), and the fused block epilogue ``scale`` / ``shift`` / ``residual`` / ``relu`` / ``want_stats``."""
    return _forward(features, weights, topology, "gatherScatterDefaultSparseConv", False, bias, **fused)


def gs_conv_backward(grad_output, features, weights, topology, on_grad_weights=None, **extra):
    """(grad_features, grad_weights) of the forward convolution (Bindings.cpp:595-609).  ``on_grad_weights(grad_weights)``
    (extension) is called as soon as the weight gradient is enqueued, before the dgrad kernel."""
    return _backward(grad_output, features, weights, topology, "gatherScatterDefaultSparseConvBackward", False, on_grad_weights, **extra)


def gs_conv_transpose(features, weights, topology, bias=None, **fused):
    """Forward transposed sparse convolution (Bindings.cpp:627-637)."""
    return _forward(features, weights, topology, "gatherScatterDefaultSparseConvTranspose", True, bias, **fused)


def gs_conv_transpose_backward(grad_output, features, weights, topology, on_grad_weights=None, **extra):
    """(grad_features, grad_weights) of the transposed convolution (Bindings.cpp:638-653)."""
    return _backward(grad_output, features, weights, topology, "gatherScatterDefaultSparseConvTransposeBackward", True, on_grad_weights, **extra)


def pred_gather_igemm_conv(features, weights, feature_grid, output_grid, kernel_size: int, stride: int):
    """Tensor-core forward (Bindings.cpp:657-674 -> predGatherIGemmSparseConv, PredGatherIGemm.cu:1121-1172).

    The reference's SM80 TF32 leaf-blocked kernel is replaced by the same tcgen05 implicit-GEMM engine that
    serves ``gs_conv``; this entry point keeps the reference's admission checks and builds the map on the fly.
    """
    if features.dim() != 2 or weights.dim() != 5:
        raise RuntimeError("predGatherIGemmSparseConv: features must be 2D and weights 5D")
    if features.size(1) % 32 or weights.size(0) % 32:
        raise RuntimeError("predGatherIGemmSparseConv requires channel counts divisible by 32")
    if feature_grid.grid_count != 1 or output_grid.grid_count != 1:
        raise RuntimeError("predGatherIGemmSparseConv supports only batch size 1")
    topology = gs_build_topology(feature_grid, output_grid, [kernel_size] * 3, [stride] * 3)
    return gs_conv(features, weights, topology)
