"""Retrieval models (public names / signatures of handyrec.models.retrieval): DSSM, YouTubeMatchDNN.

Reference behaviour reproduced (paths under /root/reference/handyrec/models/retrieval/):
  DSSM             DSSM.py:75-122            user tower over the user group (SUM pooling); item tower over `get_embd(compress=False)` of the
                                             whole catalogue; optional cosine scaling `tf.nn.l2_normalize(.)` WITHOUT an axis (norm over the
                                             whole matrix) times `gamma` on the item side; SampledSoftmaxLayer over [items, users, item id]
  YouTubeMatchDNN  YouTubeMatchDNN.py:64-100 user DNN over the user group (mean pooling); items = `get_embd(compress=True)`
Both return a Model over `user inputs + [item id]` carrying `user_input / user_embedding / item_input / item_embedding`.

`lazy_catalogue` (extra keyword, default None = "when it is exact"): evaluate the item side only for the `B + num_sampled` rows a
training step reads (`EmbdFeatureGroup.row_tower`) instead of all n items every step.  Exact unless the item side carries batch-wide
statistics (BatchNorm over the catalogue, or DSSM's whole-matrix l2 normalisation); `item_embedding` for inference always uses the
full catalogue like the reference.
"""
from __future__ import annotations

from typing import Optional, Tuple

from ..autograd_ops import L2NormalizeFn
from ..features import EmbdFeatureGroup, FeatureGroup
from ..keras_lite import Lambda, Model
from ..layers import DNN, SampledSoftmaxLayer
from ..layers.utils import concat
from ._blocks import attach_retrieval_handles, pooled_inputs, require_catalogue, rows_of, tower


def _l2_normalize(x):
    return Lambda(lambda t: L2NormalizeFn.apply(t, 1e-12), lambda s: s, name="l2_normalize")(x)


def _n_items(group: EmbdFeatureGroup) -> int:
    return len(group._value_dict[group.id_name])


def DSSM(user_feature_group: FeatureGroup, item_feature_group: EmbdFeatureGroup, user_dnn_hidden_units: Tuple[int] = (64, 32),
         item_dnn_hidden_units: Tuple[int] = (64, 32), dnn_activation: str = "relu", l2_dnn: float = 0, dnn_bn: bool = False,
         dnn_dropout: float = 0, num_sampled: int = 1, seed: int = 2022, cos_sim: bool = False, gamma: float = 10,
         lazy_catalogue: Optional[bool] = None) -> Model:
    require_catalogue(item_feature_group)
    item_id = item_feature_group.id_input
    u_dense, u_embds = pooled_inputs([user_feature_group], pool_method="sum")
    user_vec = tower(concat(u_dense, u_embds), user_dnn_hidden_units, dnn_activation, l2_dnn, dnn_dropout, dnn_bn, seed, "linear",
                     name="User_Embedding_DNN")
    item_dnn = DNN(hidden_units=tuple(item_dnn_hidden_units), activation=dnn_activation, l2_reg=l2_dnn, dropout_rate=dnn_dropout, use_bn=dnn_bn,
                   output_activation="linear", seed=seed, name="Item_Embedding_DNN")
    item_mat = item_dnn(item_feature_group.get_embd(item_id, compress=False))   # (n, d): every catalogue row
    lazy = (not cos_sim and not dnn_bn) if lazy_catalogue is None else bool(lazy_catalogue)
    if lazy:
        loss = SampledSoftmaxLayer(num_sampled=num_sampled, item_rows=item_feature_group.row_tower(compress=False, head=item_dnn),
                                   num_classes=_n_items(item_feature_group))([user_vec, item_id])
    else:
        items, users = (_l2_normalize(item_mat) * gamma, _l2_normalize(user_vec)) if cos_sim else (item_mat, user_vec)
        loss = SampledSoftmaxLayer(num_sampled=num_sampled)([items, users, item_id])
    user_inputs = list(user_feature_group.input_layers.values())
    model = Model(inputs=user_inputs + [item_id], outputs=loss)
    return attach_retrieval_handles(model, user_inputs, user_vec, item_id, rows_of(item_mat, item_id, "item_embedding"))


def YouTubeMatchDNN(user_feature_group: FeatureGroup, item_feature_group: EmbdFeatureGroup, dnn_hidden_units: Tuple[int] = (64, 32),
                    dnn_activation: str = "relu", dnn_dropout: float = 0, l2_dnn: float = 0, dnn_bn: bool = False, num_sampled: int = 1,
                    seed: int = 2022, lazy_catalogue: Optional[bool] = None) -> Model:
    require_catalogue(item_feature_group)
    item_id = item_feature_group.id_input
    u_dense, u_embds = pooled_inputs([user_feature_group], pool_method="mean")
    user_vec = tower(concat(u_dense, u_embds), dnn_hidden_units, dnn_activation, l2_dnn, dnn_dropout, dnn_bn, seed, "linear", name="User_DNN")
    item_mat = item_feature_group.get_embd(item_id, compress=True)              # (n, embd_dim): no batch statistics on the item side
    if lazy_catalogue is None or lazy_catalogue:
        loss = SampledSoftmaxLayer(num_sampled=num_sampled, item_rows=item_feature_group.row_tower(compress=True),
                                   num_classes=_n_items(item_feature_group))([user_vec, item_id])
    else:
        loss = SampledSoftmaxLayer(num_sampled=num_sampled)([item_mat, user_vec, item_id])
    user_inputs = list(user_feature_group.input_layers.values())
    model = Model(inputs=user_inputs + [item_id], outputs=loss)
    return attach_retrieval_handles(model, user_inputs, user_vec, item_id, rows_of(item_mat, item_id, "item_embedding"))
