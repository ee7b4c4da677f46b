"""Sequence layers (mirror of /root/reference/handyrec/layers/sequence.py:7-143)."""
from __future__ import annotations

import torch

from ..autograd_ops import AttInputFn, MaskScoresFn, SeqPoolFn
from ..keras_lite import Layer
from .core import DNN


class SequencePoolingLayer(Layer):
    """Masked mean / max / sum pooling over axis 1 (sequence.py:26-46)."""

    def __init__(self, method: str, **kwargs):
        super().__init__(**kwargs)
        assert method in ["mean", "max", "sum"], "Pooling method should be `mean`, `max`, or `sum`"
        self.method = method

    def call(self, inputs, mask=None):
        if mask is None:
            raise ValueError("Embedding layer should set `mask_zero` as True")
        return SeqPoolFn.apply(inputs, mask, self.method)

    def compute_output_shape(self, input_shape):
        return (None, 1, input_shape[-1])

    def get_config(self):
        return {"method": self.method, **super().get_config()}


class LocalActivationUnit(Layer):
    """The LocalActivationUnit used in DIN (sequence.py:57-143)."""

    def __init__(self, hidden_units=(64, 32, 1), activation="sigmoid", l2_reg=0, dropout_rate=0, use_bn=False, seed=1024, **kwargs):
        self.hidden_units, self.activation, self.l2_reg = hidden_units, activation, l2_reg
        self.dropout_rate, self.use_bn, self.seed = dropout_rate, use_bn, seed
        self.dnn = None
        super().__init__(**kwargs)

    def build(self, input_shape):
        self._input_check(input_shape)
        self.dnn = self._track(DNN(self.hidden_units, self.activation, self.l2_reg, self.dropout_rate, self.use_bn, seed=self.seed))
        d = input_shape[0][-1]
        self.dnn.build((None, input_shape[1][1], 4 * d))
        self.built = True

    def call(self, inputs, mask=None, **kwargs):
        query, keys = inputs  # (?, 1, D), (?, T, D)
        key_mask = mask[1]    # (?, T) after SqueezeMask
        B, T, D = keys.shape
        att_input = AttInputFn.apply(query.reshape(B, D), keys)                  # sequence.py:96-97
        att_out = self.dnn(att_input, training=kwargs.get("training", False))     # (?, T, 1)
        att_out = MaskScoresFn.apply(att_out.reshape(B, T), key_mask)             # sequence.py:100-101
        return att_out.reshape(B, 1, T)

    def compute_output_shape(self, input_shape):
        return input_shape[1][0], 1, input_shape[1][1]

    def compute_mask(self, inputs, mask):
        return mask

    def get_config(self):
        return {"activation": self.activation, "hidden_units": self.hidden_units, "l2_reg": self.l2_reg, "dropout_rate": self.dropout_rate,
                "use_bn": self.use_bn, "seed": self.seed, **super().get_config()}

    @staticmethod
    def _input_check(input_shape):
        if not isinstance(input_shape, list) or len(input_shape) != 2:
            raise ValueError("A `LocalActivationUnit` layer should be called on a list of 2 inputs")
        if len(input_shape[0]) != 3 or len(input_shape[1]) != 3:
            raise ValueError("Unexpected inputs dimensions %d and %d, expect to be 3 dimensions" % (len(input_shape[0]), len(input_shape[1])))
        if input_shape[0][-1] != input_shape[1][-1] or input_shape[0][1] != 1:
            raise ValueError("A `LocalActivationUnit` layer requires inputs of a two inputs with shape (None,1,embedding_size) and "
                             "(None,T,embedding_size) Got different shapes: %s,%s" % (input_shape[0], input_shape[1]))
