from .modules import AvgPool, BatchNorm, MaxPool, SparseConv3d, SparseConvTranspose3d, SyncBatchNorm, UpsamplingNearest

__all__ = ["AvgPool", "BatchNorm", "MaxPool", "SparseConv3d", "SparseConvTranspose3d", "SyncBatchNorm", "UpsamplingNearest"]
