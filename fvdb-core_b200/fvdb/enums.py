"""Convolution enums (mirror of reference fvdb/enums.py:16-43; values are part of the public API)."""

from enum import StrEnum


class ConvolutionTopologyPolicy(StrEnum):
    """Policy controlling the finite output topology of a convolution plan."""

    COMPLETE = "complete"
    RESTRICTED = "restricted"


class ConvolutionTopologyProvenance(StrEnum):
    """How a convolution plan's finite topology was obtained."""

    GENERATED = "generated"
    EXPLICIT_TARGET = "explicit_target"
    EXACT_TRANSPOSE = "exact_transpose"


class ConvolutionPhasePolicy(StrEnum):
    """Kernel phase convention used by a convolution plan."""

    TORCH_SAME_PHASE = "torch_same_phase"
