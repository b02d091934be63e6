"""``GridBatch``: the subset of the reference's grid API the convolution path uses.

Mirrors the names and meaning of reference fvdb/grid_batch.py (``from_ijk:220``, ``conv_grid:582``,
``conv_transpose_grid:606``, ``jagged_like:1026``, ``neighbor_indexes:1181``, ``ijk:1944``, ``grid_count:1926``,
``total_voxels:2084``, ``is_same:1013``, ``voxel_sizes`` / ``origins``).  Everything else in the reference's
~100-method grid API (sampling, rays, meshing, IO ...) is outside the ConvolutionPlan path.
Grids live on a CUDA device; there is no CPU grid in this build.
"""

from __future__ import annotations

import torch

from . import _fvdb_cpp
from .jagged_tensor import JaggedTensor
from .types import NumericMaxRank1, NumericMaxRank2, ValueConstraint, to_Vec3fBatch, to_Vec3i


class GridBatch:
    def __init__(self, data: _fvdb_cpp.GridBatchData):
        if not isinstance(data, _fvdb_cpp.GridBatchData):
            raise TypeError("GridBatch wraps a GridBatchData; use GridBatch.from_ijk(...)")
        self.data = data

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def from_ijk(cls, ijk: "JaggedTensor | torch.Tensor", voxel_sizes: NumericMaxRank2 = 1, origins: NumericMaxRank2 = 0) -> "GridBatch":
        """Grid batch from explicit voxel coordinates (duplicates and negative coordinates allowed)."""
        if isinstance(ijk, torch.Tensor):
            ijk = JaggedTensor(ijk)
        num_grids = ijk.num_tensors
        sizes = to_Vec3fBatch(voxel_sizes, num_grids, "voxel_sizes", positive=True)
        orig = to_Vec3fBatch(origins, num_grids, "origins")
        coords = ijk.jdata.reshape(-1, 3) if ijk.jdata.numel() == 0 else ijk.jdata
        data = _fvdb_cpp.build_grid_from_ijk(coords, ijk.jidx if num_grids > 1 else None, num_grids, sizes, orig)
        return cls(data)

    @classmethod
    def from_points(cls, points: "JaggedTensor | torch.Tensor", voxel_sizes: NumericMaxRank2 = 1, origins: NumericMaxRank2 = 0) -> "GridBatch":
        """Grid batch with one voxel per occupied point location (fvdb/grid_batch.py:389-409).

        ``ijk = round(p / voxel_size - origin / voxel_size)`` in the points' own precision, exactly the primal
        transform of src/fvdb/VoxelCoordTransform.h:300-310 applied by ops/BuildGridFromPoints.cu:89,119-124;
        ``round`` is NanoVDB's ``Vec3::round`` = ``floor(x + 0.5)`` (recalled upstream convention: NanoVDB is not
        vendored in the reference tree).  The coordinates then go through the same device builder as ``from_ijk``."""
        if isinstance(points, torch.Tensor):
            points = JaggedTensor(points)
        if not points.jdata.is_floating_point():
            raise TypeError("points must be a floating point JaggedTensor")
        num_grids = points.num_tensors
        sizes = to_Vec3fBatch(voxel_sizes, num_grids, "voxel_sizes", positive=True)
        orig = to_Vec3fBatch(origins, num_grids, "origins")
        pts = points.jdata.reshape(-1, 3)
        math_dtype = torch.float64 if pts.dtype == torch.float64 else torch.float32
        scale = (1.0 / sizes.double()).to(device=pts.device, dtype=math_dtype)
        shift = (-orig.double() / sizes.double()).to(device=pts.device, dtype=math_dtype)
        which = points.jidx.long() if num_grids > 1 else torch.zeros(pts.shape[0], dtype=torch.long, device=pts.device)
        ijk = torch.floor(pts.to(math_dtype) * scale[which] + shift[which] + 0.5).to(torch.int32)
        data = _fvdb_cpp.build_grid_from_ijk(ijk, points.jidx if num_grids > 1 else None, num_grids, sizes, orig)
        return cls(data)

    @classmethod
    def from_zero_voxels(cls, device="cuda", voxel_sizes: NumericMaxRank2 = 1, origins: NumericMaxRank2 = 0) -> "GridBatch":
        sizes = torch.as_tensor(voxel_sizes, dtype=torch.float64).reshape(-1, 3) if not isinstance(voxel_sizes, (int, float)) else None
        num_grids = 1 if sizes is None else int(sizes.shape[0])
        empty = [torch.zeros((0, 3), dtype=torch.int32, device=device) for _ in range(num_grids)]
        return cls.from_ijk(JaggedTensor(empty), voxel_sizes, origins)

    # ---- metadata -------------------------------------------------------------------------
    @property
    def device(self) -> torch.device:
        return self.data.device

    @property
    def grid_count(self) -> int:
        return self.data.num_grids

    def __len__(self) -> int:
        return self.data.num_grids

    @property
    def total_voxels(self) -> int:
        return self.data.total_voxels

    @property
    def total_leaf_nodes(self) -> int:
        return self.data.num_leaves

    @property
    def num_voxels(self) -> torch.Tensor:
        offsets = self.data.voxel_offsets
        return offsets[1:] - offsets[:-1]

    def num_voxels_at(self, bi: int) -> int:
        return int(self.num_voxels[bi].item())

    @property
    def joffsets(self) -> torch.Tensor:
        return self.data.voxel_offsets

    @property
    def jidx(self) -> torch.Tensor:
        return self.data.jidx

    @property
    def voxel_sizes(self) -> torch.Tensor:
        return self.data.voxel_sizes.to(device=self.device, dtype=torch.float32)

    @property
    def origins(self) -> torch.Tensor:
        return self.data.origins.to(device=self.device, dtype=torch.float32)

    @property
    def ijk(self) -> JaggedTensor:
        """Voxel coordinates in row order, one tensor per grid."""
        return JaggedTensor(_data=self.data.ijk, _offsets=self.data.voxel_offsets, _jidx=self.data.jidx)

    def is_same(self, other: "GridBatch") -> bool:
        return self.data.is_same(other.data)

    def jagged_like(self, data: torch.Tensor) -> JaggedTensor:
        """Wrap ``data [total_voxels, ...]`` with this batch's per-grid structure."""
        if data.shape[0] != self.total_voxels:
            raise ValueError(f"data has {data.shape[0]} rows but the grid batch has {self.total_voxels} voxels")
        jidx = self.data.jidx
        return JaggedTensor(_data=data, _offsets=self.data.voxel_offsets.to(data.device), _jidx=jidx.to(data.device))

    # ---- generated convolution topologies --------------------------------------------------
    def conv_grid(self, kernel_size: NumericMaxRank1, stride: NumericMaxRank1 = 1) -> "GridBatch":
        """Complete structural support of a convolution on this batch (voxel size * stride, same origin).
        ``kernel_size = stride = 1`` returns ``self`` (reference functional/_topology.py:452-453)."""
        ks = to_Vec3i(kernel_size, value_constraint=ValueConstraint.POSITIVE).tolist()
        st = to_Vec3i(stride, value_constraint=ValueConstraint.POSITIVE).tolist()
        if ks == [1, 1, 1] and st == [1, 1, 1]:
            return self
        return GridBatch(_fvdb_cpp.conv_grid(self.data, ks, st))

    def conv_transpose_grid(self, kernel_size: NumericMaxRank1, stride: NumericMaxRank1 = 1) -> "GridBatch":
        """Complete structural support of a transposed convolution (voxel size / stride, same origin)."""
        ks = to_Vec3i(kernel_size, value_constraint=ValueConstraint.POSITIVE).tolist()
        st = to_Vec3i(stride, value_constraint=ValueConstraint.POSITIVE).tolist()
        if ks == [1, 1, 1] and st == [1, 1, 1]:
            return self
        return GridBatch(_fvdb_cpp.conv_transpose_grid(self.data, ks, st))

    # ---- coarsening / refinement (block-centroid transforms, not the convolution lattice) --------------------------
    def _derived(self, ijk: torch.Tensor, jidx: "torch.Tensor | None", sizes: torch.Tensor, origins: torch.Tensor) -> "GridBatch":
        return GridBatch(_fvdb_cpp.build_grid_from_ijk(ijk, jidx if self.grid_count > 1 else None, self.grid_count, sizes, origins))

    def coarsened_grid(self, coarsening_factor: NumericMaxRank1) -> "GridBatch":
        """Coarse voxels ``floor(ijk / factor)`` (ops/BuildCoarseGridFromFine.cu:50,136); voxel size ``* factor``, origin
        ``+ (factor - 1) * voxel_size / 2`` (detail/utils/VoxelSizeUtils.h:13-22)."""
        f = to_Vec3i(coarsening_factor, value_constraint=ValueConstraint.POSITIVE)
        fd = f.to(torch.float64)
        sizes, origins = self.data.voxel_sizes.double(), self.data.origins.double()
        coarse = torch.div(self.data.ijk, f.to(self.device, torch.int32), rounding_mode="floor").to(torch.int32)
        return self._derived(coarse, self.data.jidx, sizes * fd, (fd - 1.0) * sizes * 0.5 + origins)

    def refined_grid(self, subdiv_factor: NumericMaxRank1, mask: "JaggedTensor | torch.Tensor | None" = None) -> "GridBatch":
        """Fine voxels ``factor * ijk + [0, factor)^3`` of every (selected) voxel (ops/BuildFineGridFromCoarse.cu); voxel size
        ``/ factor``, origin ``- (factor - 1) * fine_voxel_size / 2`` (VoxelSizeUtils.h:24-35)."""
        f = to_Vec3i(subdiv_factor, value_constraint=ValueConstraint.POSITIVE)
        fd = f.to(torch.float64)
        sizes, origins = self.data.voxel_sizes.double(), self.data.origins.double()
        ijk, jidx = self.data.ijk, self.data.jidx
        if mask is not None:
            keep = (mask.jdata if isinstance(mask, JaggedTensor) else mask).to(torch.bool).reshape(-1)
            ijk, jidx = ijk[keep], jidx[keep]
        f0, f1, f2 = f.tolist()
        cells = torch.stack(torch.meshgrid(torch.arange(f0), torch.arange(f1), torch.arange(f2), indexing="ij"), dim=-1).reshape(-1, 3).to(self.device, torch.int32)
        fine = (ijk[:, None, :] * f.to(self.device, torch.int32) + cells[None]).reshape(-1, 3).contiguous()
        return self._derived(fine, jidx.repeat_interleave(cells.shape[0]), sizes / fd, origins - (fd - 1.0) * (sizes / fd) * 0.5)

    def _pool(self, mode: int, pool_factor, data: JaggedTensor, stride, coarse_grid: "GridBatch | None"):
        from . import _pool

        factor = to_Vec3i(pool_factor, value_constraint=ValueConstraint.POSITIVE).tolist()
        st = to_Vec3i(stride, value_constraint=ValueConstraint.NON_NEGATIVE).tolist()
        st = [s if s > 0 else f for s, f in zip(st, factor)]  # stride 0 = pool_factor (MaxPool.cu:137-140)
        if any(s < f for s, f in zip(st, factor)):
            raise ValueError("pooling windows must not overlap: stride >= pool_factor on every axis")
        if coarse_grid is None:
            coarse_grid = self.coarsened_grid(st)
        if data.jdata.shape[0] != self.total_voxels:
            raise ValueError("data must have one row per voxel of the fine grid")
        idx = _pool.window_children(self, coarse_grid, factor, st)
        scale = 1.0 if mode == _pool.POOL_MAX else 1.0 / (factor[0] * factor[1] * factor[2])  # AvgPool.cu:145: over the whole window
        out = _pool.pool_rows(data.jdata, idx, mode, scale)
        return coarse_grid.jagged_like(out), coarse_grid

    def max_pool(self, pool_factor: NumericMaxRank1, data: JaggedTensor, stride: NumericMaxRank1 = 0, coarse_grid: "GridBatch | None" = None):
        """``(pooled, coarse_grid)``: channel-wise max over the active voxels of each window (fvdb/grid_batch.py:1139-1164)."""
        from . import _pool

        return self._pool(_pool.POOL_MAX, pool_factor, data, stride, coarse_grid)

    def avg_pool(self, pool_factor: NumericMaxRank1, data: JaggedTensor, stride: NumericMaxRank1 = 0, coarse_grid: "GridBatch | None" = None):
        """``(pooled, coarse_grid)``: sum over the active voxels of each window divided by the window volume (grid_batch.py:463-490)."""
        from . import _pool

        return self._pool(_pool.POOL_SUM, pool_factor, data, stride, coarse_grid)

    def refine(self, subdiv_factor: NumericMaxRank1, data: JaggedTensor, mask: "JaggedTensor | None" = None, fine_grid: "GridBatch | None" = None):
        """``(refined, fine_grid)``: every fine voxel takes the features of ``floor(ijk / factor)`` (grid_batch.py:1474-1499)."""
        from . import _pool

        factor = to_Vec3i(subdiv_factor, value_constraint=ValueConstraint.POSITIVE).tolist()
        if fine_grid is None:
            fine_grid = self.refined_grid(factor, mask)
        if data.jdata.shape[0] != self.total_voxels:
            raise ValueError("data must have one row per voxel of the coarse grid")
        parent = _pool.parent_rows(self, fine_grid, factor)
        children = _pool.window_children(fine_grid, self, factor, factor)
        out = _pool.refine_rows(data.jdata, parent, children)
        return fine_grid.jagged_like(out), fine_grid

    # ---- lookups ----------------------------------------------------------------------------
    def _query(self, ijk: "JaggedTensor | torch.Tensor") -> JaggedTensor:
        if isinstance(ijk, torch.Tensor):
            ijk = JaggedTensor(ijk)
        if ijk.num_tensors != self.grid_count:
            raise ValueError(f"query has {ijk.num_tensors} tensors but the grid batch has {self.grid_count} grids")
        return ijk

    def neighbor_indexes(self, ijk: "JaggedTensor | torch.Tensor", extent: int, bitshift: int = 0) -> JaggedTensor:
        """Per-grid-local index (or -1) of every voxel in ``[-extent, extent]^3`` around each query."""
        q = self._query(ijk)
        out = _fvdb_cpp.neighbor_indexes(self.data, q.jdata, q.jidx if self.grid_count > 1 else None, extent, bitshift)
        return q.jagged_like(out)

    def ijk_to_index(self, ijk: "JaggedTensor | torch.Tensor", cumulative: bool = False) -> JaggedTensor:
        q = self._query(ijk)
        out = _fvdb_cpp.ijk_to_index(self.data, q.jdata, q.jidx if self.grid_count > 1 else None, cumulative)
        return q.jagged_like(out)

    def coords_in_grid(self, ijk: "JaggedTensor | torch.Tensor") -> JaggedTensor:
        idx = self.ijk_to_index(ijk)
        return idx.jagged_like(idx.jdata >= 0)

    def __repr__(self) -> str:
        return f"GridBatch(grid_count={self.grid_count}, total_voxels={self.total_voxels}, device={self.device})"
