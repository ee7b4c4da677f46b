"""DeepFM (mirror of /root/reference/handyrec/models/ranking/context_aware/DeepFM.py:10-94)."""
from __future__ import annotations

import warnings
from typing import Tuple

from ..features import FeatureGroup
from ..keras_lite import Activation, Model
from ..layers import DNN, FM
from ..layers.utils import concat


def DeepFM(fm_feature_group: FeatureGroup, dnn_feature_group: FeatureGroup, dnn_hidden_units: Tuple[int] = (64, 32, 1),
           dnn_activation: str = "relu", dnn_dropout: float = 0, dnn_bn: bool = False, l2_dnn: float = 0, task: str = "binary",
           seed: int = 2022) -> Model:
    if dnn_hidden_units[-1] != 1:  # DeepFM.py:59-60
        raise ValueError("Output size of dnn should be 1")
    fm_dense, fm_sparse = fm_feature_group.embedding_lookup(pool_method="mean")
    dnn_dense, dnn_sparse = dnn_feature_group.embedding_lookup(pool_method="mean")
    if len(fm_dense) > 0:
        warnings.warn("FM currently doesn't support dense featrue, they will be ignored")
    dnn_input = concat(dnn_dense, dnn_sparse)                       # (B, sum dense + sum D), dense first
    fm_input = concat([], fm_sparse, axis=1, keepdims=True)         # (B, F, D)
    dnn_output = DNN(hidden_units=dnn_hidden_units, activation=dnn_activation, l2_reg=l2_dnn, dropout_rate=dnn_dropout, use_bn=dnn_bn,
                     output_activation="linear", seed=seed, name="Deep_Part")(dnn_input)
    fm_output = FM(name="FM_Part")(fm_input)
    output = dnn_output + fm_output                                 # DeepFM.py:86
    if task == "binary":
        output = Activation("sigmoid")(output)
    inputs = list(fm_feature_group.feat_pool.input_layers.values())
    return Model(inputs=inputs, outputs=output)
