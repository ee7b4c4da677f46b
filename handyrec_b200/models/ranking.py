"""Ranking models (public names / signatures of handyrec.models.ranking): DeepFM, DIN, YouTubeRankDNN.

Reference behaviour reproduced (paths under /root/reference/handyrec/models/ranking/):
  DeepFM          context_aware/DeepFM.py:59-94   both groups looked up with mean pooling; DNN(concat(dense, embeddings)) with a
                                                  linear 1-unit head + FM(embeddings stacked on axis 1); sigmoid for task="binary";
                                                  dense features of the FM group are ignored with a warning; inputs = every input of the pool
  DIN             sequential/DIN.py:66-112        per behaviour sequence: LocalActivationUnit scores (masked, no softmax) x keys;
                                                  DNN over [other dense | other embeddings | attention pools] with an extra sigmoid unit
  YouTubeRankDNN  context_aware/YouTubeRankDNN.py:49-69  DNN over user + item groups with an extra sigmoid unit
Graphs shaped like DeepFM are recognised by `handyrec_b200.lowering` at `compile()` and trained on the fused engine.
"""
from __future__ import annotations

import warnings
from typing import Tuple

from ..features import FeatureGroup
from ..keras_lite import Activation, Model
from ..layers import FM, AttentionPooling, LocalActivationUnit, SqueezeMask
from ..layers.utils import concat
from ._blocks import all_inputs, pooled_inputs, tower


def DeepFM(fm_feature_group: FeatureGroup, dnn_feature_group: FeatureGroup, dnn_hidden_units: Tuple[int] = (64, 32, 1),
           dnn_activation: str = "relu", dnn_dropout: float = 0, dnn_bn: bool = False, l2_dnn: float = 0, task: str = "binary",
           seed: int = 2022) -> Model:
    if dnn_hidden_units[-1] != 1:
        raise ValueError("Output size of dnn should be 1")
    fm_dense, fm_embds = pooled_inputs([fm_feature_group])
    deep_dense, deep_embds = pooled_inputs([dnn_feature_group])
    if fm_dense:
        warnings.warn("FM currently doesn't support dense featrue, they will be ignored")
    deep = tower(concat(deep_dense, deep_embds), dnn_hidden_units, dnn_activation, l2_dnn, dnn_dropout, dnn_bn, seed, "linear", name="Deep_Part")
    wide = FM(name="FM_Part")(concat([], fm_embds, axis=1, keepdims=True))
    score = deep + wide
    if task == "binary":
        score = Activation("sigmoid")(score)
    return Model(inputs=all_inputs(fm_feature_group.feat_pool), outputs=score)


def DIN(item_seq_feat_group: FeatureGroup, other_feature_group: FeatureGroup, dnn_hidden_units: Tuple[int] = (64, 32, 1),
        dnn_activation: str = "dice", dnn_dropout: float = 0, dnn_bn: bool = False, l2_dnn: float = 0,
        lau_dnn_hidden_units: Tuple[int] = (32, 1), lau_dnn_activation: str = "dice", lau_dnn_dropout: float = 0, lau_dnn_bn: bool = False,
        lau_l2_dnn: float = 0, seed: int = 2022) -> Model:
    pool = item_seq_feat_group.feat_pool
    ctx_dense, ctx_embds = pooled_inputs([other_feature_group])
    interests = []
    target = None
    for seq in item_seq_feat_group.features:
        if target is None:  # the candidate item: one (B,1) input named after the unit of the FIRST sequence feature (DIN.py:72-77)
            target = pool.init_input(seq.unit.name, {"name": seq.unit.name, "shape": (1,), "dtype": "int32"})
        table = item_seq_feat_group.embd_layers[seq.unit.name]
        keys = SqueezeMask()(table(item_seq_feat_group.input_layers[seq.name]))   # (B,T,D), mask (B,T)
        scores = LocalActivationUnit(lau_dnn_hidden_units, lau_dnn_activation, lau_l2_dnn, lau_dnn_dropout, lau_dnn_bn, seed)([table(target), keys])
        interests.append(AttentionPooling()([scores, keys]))                        # tf.matmul(att_score, embd_seq) -> (B,1,D)
    out = tower(concat(ctx_dense, ctx_embds + interests), tuple(dnn_hidden_units) + (1,), dnn_activation, l2_dnn, dnn_dropout, dnn_bn, seed, "sigmoid")
    return Model(inputs=all_inputs(pool), outputs=out)


def YouTubeRankDNN(user_feature_group: FeatureGroup, item_feature_group: FeatureGroup, dnn_hidden_units: Tuple[int] = (64, 32),
                   dnn_activation: str = "relu", dnn_dropout: float = 0, l2_dnn: float = 0, dnn_bn: bool = False, seed: int = 2022) -> Model:
    # dense inputs of BOTH groups come first, then the embeddings of both (YouTubeRankDNN.py:53)
    u_dense, u_embds = pooled_inputs([user_feature_group])
    i_dense, i_embds = pooled_inputs([item_feature_group])
    out = tower(concat(u_dense + i_dense, u_embds + i_embds), tuple(dnn_hidden_units) + (1,), dnn_activation, l2_dnn, dnn_dropout, dnn_bn, seed,
                "sigmoid")
    inputs = list(user_feature_group.input_layers.values()) + list(item_feature_group.input_layers.values())
    return Model(inputs=inputs, outputs=out)
