"""Layer utilities (mirror of /root/reference/handyrec/layers/utils.py:9-133)."""
from __future__ import annotations

from typing import Any, List

import numpy as np
import torch

from ..keras_lite import KTensor, Lambda, Layer
from .core import get_activation_layer  # noqa: F401  (same import location as the reference)


class Concatenate(Layer):
    """keras Concatenate.  (The zero-copy form of this concat is the fused engine's layout contract, DESIGN.md 2: `lowering.py` puts
    DeepFM-shaped graphs there; on the layer path it is a plain copy.)"""

    def __init__(self, axis=-1, **kw):
        super().__init__(**kw)
        self.axis = axis

    def call(self, inputs):
        return torch.cat(list(inputs), dim=self.axis)

    def compute_output_shape(self, input_shape):
        ax = self.axis if self.axis >= 0 else len(input_shape[0]) + self.axis
        out = list(input_shape[0])
        out[ax] = sum(s[ax] for s in input_shape)
        return tuple(out)

    def output_dtype(self, inputs):
        return inputs[0].dtype.name


class Flatten(Layer):
    def call(self, inputs):
        return inputs.reshape(inputs.shape[0], -1)

    def compute_output_shape(self, input_shape):
        n = 1
        for s in input_shape[1:]:
            n *= s
        return (input_shape[0], n)

    def output_dtype(self, inputs):
        return inputs.dtype.name


class Cast(Layer):
    def call(self, inputs):
        return inputs.to(torch.float32)


def _concat(inputs: List, axis: int = -1):
    if len(inputs) == 1:  # utils.py:24-25
        return inputs[0]
    has_integer = any(t.dtype.is_integer if isinstance(t, KTensor) else not t.dtype.is_floating_point for t in inputs)
    has_other = any(not (t.dtype.is_integer if isinstance(t, KTensor) else not t.dtype.is_floating_point) for t in inputs)
    if has_other and has_integer:  # utils.py:28-36
        inputs = [Cast()(t) for t in inputs]
    return Concatenate(axis=axis)(inputs)


def concat(dense_inputs: List, embd_inputs: List, axis: int = -1, keepdims: bool = False):
    """Concatenate dense features and embeddings of sparse features (utils.py:40-96): dense part first."""
    if len(dense_inputs) + len(embd_inputs) == 0:
        raise ValueError("Number of inputs should be larger than 0")
    if len(dense_inputs) > 0 and len(embd_inputs) > 0:
        dense = _concat(dense_inputs, axis)
        sparse = _concat(embd_inputs, axis)
        if not keepdims:
            dense = Flatten()(dense)
            sparse = Flatten()(sparse)
        return _concat([dense, sparse], axis)
    output = _concat(dense_inputs if len(dense_inputs) > 0 else embd_inputs, axis)
    if not keepdims:
        output = Flatten()(output)
    return output


def sampledsoftmaxloss(y_true, y_pred) -> Any:
    """utils.py:99-114: `tf.reduce_mean(y_pred)`."""
    if isinstance(y_pred, torch.Tensor):
        return y_pred.mean()
    return float(np.mean(np.asarray(y_pred, dtype=np.float64)))
