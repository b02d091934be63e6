from .modules import SparseConv3d, SparseConvTranspose3d

__all__ = ["SparseConv3d", "SparseConvTranspose3d"]
