"""Tool layers (mirror of /root/reference/handyrec/layers/tools.py)."""
from __future__ import annotations

from typing import List

import numpy as np
import torch

from ..autograd_ops import AttPoolFn, EmbeddingFn
from ..keras_lite import KTensor, Layer, device


class ValueTable(Layer):
    """Output the full list of values of a feature (tools.py:10-29): `call` ignores its input."""

    def __init__(self, value_list: List, dtype: str = "int32", **kwargs):
        self._np = np.asarray(value_list)
        self._dtype = dtype
        self.value = None
        super().__init__(**kwargs)

    def call(self, *args, **kwargs):
        if self.value is None:
            from ..keras_lite import _tdtype

            self.value = torch.as_tensor(self._np).to(device(), _tdtype(self._dtype))
        return self.value

    def compute_output_shape(self, input_shape):
        return tuple(self._np.shape)

    def output_dtype(self, inputs):
        return self._dtype


class CustomEmbedding(Layer):
    """Keras Embedding whose mask has the SAME shape as the output (tools.py:87-101).

    Arguments follow keras.layers.Embedding: input_dim, output_dim, embeddings_regularizer (an l2 factor or an object with
    `.l2`), trainable, weights=[ndarray], mask_zero, name.
    """

    def __init__(self, input_dim, output_dim, embeddings_regularizer=None, trainable=True, weights=None, mask_zero=False, **kwargs):
        super().__init__(trainable=trainable, **kwargs)
        self.input_dim, self.output_dim, self.mask_zero = int(input_dim), int(output_dim), bool(mask_zero)
        self.l2 = float(getattr(embeddings_regularizer, "l2", embeddings_regularizer or 0.0))
        self._init_weights = weights
        self.embeddings = None

    def build(self, input_shape):
        init = None if self._init_weights is None else np.asarray(self._init_weights[0], dtype=np.float32)
        self.embeddings = self.add_weight("embeddings", (self.input_dim, self.output_dim), initializer="uniform", l2=self.l2, value=init)
        self.built = True

    def call(self, inputs):
        ids = inputs.to(torch.int32)
        out, mask = EmbeddingFn.apply(self.embeddings, ids, self.mask_zero)
        self._last_mask = mask if self.mask_zero else None
        return out

    def compute_mask(self, inputs, mask=None):
        if not self.mask_zero:  # tools.py:94-95
            return None
        return self._last_mask  # tile(expand_dims(inputs != 0), output_dim), produced by the same kernel

    def symbolic_has_mask(self, inputs, in_masks):
        return self.mask_zero

    def compute_output_shape(self, input_shape):
        return tuple(input_shape) + (self.output_dim,)


class SqueezeMask(Layer):
    """tools.py:104-113: pass the values through, keep one mask column."""

    def call(self, inputs, *args, **kwargs):
        return inputs + 0 if not isinstance(inputs, KTensor) else inputs

    def compute_mask(self, inputs, mask=None):
        if mask is None:
            return None
        return mask[:, :, 0]

    def symbolic_has_mask(self, inputs, in_masks):
        return in_masks is not None


class AttentionPooling(Layer):
    """`tf.matmul(att_score (B,1,T), embd_seq (B,T,D))` of models/ranking/sequential/DIN.py:93 as a layer."""

    def call(self, inputs):
        att, keys = inputs
        return AttPoolFn.apply(att.reshape(att.shape[0], -1), keys)

    def compute_output_shape(self, input_shape):
        return (input_shape[1][0], 1, input_shape[1][2])
