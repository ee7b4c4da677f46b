"""DIN (mirror of /root/reference/handyrec/models/ranking/sequential/DIN.py:9-112)."""
from __future__ import annotations

from collections import OrderedDict
from typing import Tuple

from ..features import FeatureGroup
from ..keras_lite import Model
from ..layers import DNN, AttentionPooling, LocalActivationUnit, SqueezeMask
from ..layers.utils import concat


def DIN(item_seq_feat_group: FeatureGroup, other_feature_group: FeatureGroup, dnn_hidden_units: Tuple[int] = (64, 32, 1),
        dnn_activation: str = "dice", dnn_dropout: float = 0, dnn_bn: bool = False, l2_dnn: float = 0,
        lau_dnn_hidden_units: Tuple[int] = (32, 1), lau_dnn_activation: str = "dice", lau_dnn_dropout: float = 0, lau_dnn_bn: bool = False,
        lau_l2_dnn: float = 0, seed: int = 2022) -> Model:
    feature_pool = item_seq_feat_group.feat_pool
    other_dense, other_sparse = other_feature_group.embedding_lookup(pool_method="mean")
    embd_outputs = OrderedDict()
    id_input = None
    for feat in item_seq_feat_group.features:
        if id_input is None:
            id_input = feature_pool.init_input(feat.unit.name, {"name": feat.unit.name, "shape": (1,), "dtype": "int32"})
        sparse_embd = item_seq_feat_group.embd_layers[feat.unit.name]
        seq_input = item_seq_feat_group.input_layers[feat.name]
        lau = LocalActivationUnit(lau_dnn_hidden_units, lau_dnn_activation, lau_l2_dnn, lau_dnn_dropout, lau_dnn_bn, seed)
        embd_seq = sparse_embd(seq_input)            # (B, T, D) + mask (B, T, D)
        embd_seq = SqueezeMask()(embd_seq)            # mask -> (B, T)
        query = sparse_embd(id_input)                 # (B, 1, D)
        att_score = lau([query, embd_seq])            # (B, 1, T), masked, no softmax
        embd_outputs[feat.name] = AttentionPooling()([att_score, embd_seq])  # tf.matmul(att_score, embd_seq), DIN.py:93
    local_activate_pool = list(embd_outputs.values())
    dnn_input = concat(other_dense, other_sparse + local_activate_pool)
    dnn_output = DNN(hidden_units=tuple(list(dnn_hidden_units) + [1]), activation=dnn_activation, output_activation="sigmoid", l2_reg=l2_dnn,
                     dropout_rate=dnn_dropout, use_bn=dnn_bn, seed=seed)(dnn_input)
    inputs = list(feature_pool.input_layers.values())
    return Model(inputs=inputs, outputs=dnn_output)
