from .modules import BatchNorm, SparseConv3d, SparseConvTranspose3d, SyncBatchNorm

__all__ = ["BatchNorm", "SparseConv3d", "SparseConvTranspose3d", "SyncBatchNorm"]
