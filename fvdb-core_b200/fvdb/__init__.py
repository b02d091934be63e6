"""B200-native drop-in for fVDB's ConvolutionPlan path (``import fvdb`` with this directory on sys.path).

Exports the names the reference's convolution users import from ``fvdb``: ``GridBatch``, ``JaggedTensor``,
``ConvolutionPlan``, the convolution enums, ``fvdb.nn.SparseConv3d`` / ``SparseConvTranspose3d`` and the
low-level ``_fvdb_cpp`` module.  Importing fails loudly if libfvdbconv.so has not been built.
"""

from . import _fvdb_cpp, nn
from .convolution_plan import ConvolutionCoverageReport, ConvolutionCoverageWarning, ConvolutionPlan, ConvolutionTransformCompatibility
from .enums import ConvolutionPhasePolicy, ConvolutionTopologyPolicy, ConvolutionTopologyProvenance
from .grid_batch import GridBatch
from .jagged_tensor import JaggedTensor

__version__ = "0.1.0+b200"

__all__ = [
    "GridBatch",
    "JaggedTensor",
    "ConvolutionPlan",
    "ConvolutionCoverageReport",
    "ConvolutionCoverageWarning",
    "ConvolutionTransformCompatibility",
    "ConvolutionPhasePolicy",
    "ConvolutionTopologyPolicy",
    "ConvolutionTopologyProvenance",
    "nn",
    "_fvdb_cpp",
]
