from .modules import AvgPool, BatchNorm, GroupNorm, MaxPool, conv_bn_act, SparseConv3d, SparseConvTranspose3d, SyncBatchNorm, UpsamplingNearest
from .simple_unet import (
    SimpleUNet,
    SimpleUNetBasicBlock,
    SimpleUNetBottleneck,
    SimpleUNetConvBlock,
    SimpleUNetDown,
    SimpleUNetDownUp,
    SimpleUNetPad,
    SimpleUNetUnpad,
    SimpleUNetUp,
)

__all__ = [
    "AvgPool", "BatchNorm", "GroupNorm", "MaxPool", "SimpleUNet", "SimpleUNetBasicBlock", "SimpleUNetBottleneck", "SimpleUNetConvBlock", "SimpleUNetDown",
    "SimpleUNetDownUp", "SimpleUNetPad", "SimpleUNetUnpad", "SimpleUNetUp", "SparseConv3d", "SparseConvTranspose3d", "SyncBatchNorm",
    "UpsamplingNearest", "conv_bn_act",
]
