"""The slice of ``fvdb.torch_jagged`` the convolution networks use: element-wise functions lifted to JaggedTensor."""

from __future__ import annotations

import torch

from .jagged_tensor import JaggedTensor


def relu(x: JaggedTensor) -> JaggedTensor:
    return x.jagged_like(torch.relu(x.jdata))


def sigmoid(x: JaggedTensor) -> JaggedTensor:
    return x.jagged_like(torch.sigmoid(x.jdata))


def tanh(x: JaggedTensor) -> JaggedTensor:
    return x.jagged_like(torch.tanh(x.jdata))
