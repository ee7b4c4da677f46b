"""Building blocks shared by the ranking and retrieval constructors."""
from __future__ import annotations

from typing import List, Optional, Sequence

from ..features import EmbdFeatureGroup, FeatureGroup
from ..keras_lite import KTensor, Lambda, Model
from ..layers import DNN
from ..layers.utils import concat


def pooled_inputs(groups: Sequence[FeatureGroup], pool_method: str = "mean"):
    """(dense inputs, embeddings) of one or more groups, group after group -- the order the reference concatenates them in."""
    dense: List[KTensor] = []
    sparse: List[KTensor] = []
    for g in groups:
        d, s = g.embedding_lookup(pool_method=pool_method)
        dense += d
        sparse += s
    return dense, sparse


def tower(x: KTensor, units: Sequence[int], activation: str, l2: float, dropout: float, bn: bool, seed: int,
          output_activation: Optional[str], name: Optional[str] = None) -> KTensor:
    """One `DNN` (layers/core.py) applied to `x`."""
    kw = {"name": name} if name else {}
    return DNN(hidden_units=tuple(units), activation=activation, l2_reg=l2, dropout_rate=dropout, use_bn=bn,
               output_activation=output_activation, seed=seed, **kw)(x)


def all_inputs(pool) -> List[KTensor]:
    return list(pool.input_layers.values())


def require_catalogue(group) -> None:
    if not isinstance(group, EmbdFeatureGroup):
        raise ValueError("Item feature group should be an instance of `EmbdFeatureGroup`!")


def rows_of(matrix: KTensor, index: KTensor, name: str) -> KTensor:
    """`tf.squeeze(tf.nn.embedding_lookup(matrix, index), axis=1)`: rows of a computed (n, d) matrix for ids (B, 1)."""
    from ..autograd_ops import EmbeddingFn

    def gather(m, idx):
        out, _ = EmbeddingFn.apply(m, idx.reshape(-1).to(m.device).int().contiguous(), False)
        return out

    return Lambda(gather, lambda sm, si: (si[0], sm[-1]), name=name)([matrix, index])


def attach_retrieval_handles(model: Model, user_inputs, user_vec, item_id, item_vec) -> Model:
    """The four attributes the reference's inference code reads off a retrieval model (DSSM.py:116-119)."""
    model.user_input, model.user_embedding = user_inputs, user_vec
    model.item_input, model.item_embedding = item_id, item_vec
    return model
