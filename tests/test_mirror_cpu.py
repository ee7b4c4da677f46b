"""CPU: the handyrec.features / handyrec.layers mirror -- the reference's own shape / exception tests re-expressed
(/root/reference/tests/features/test_group.py, test_feature_utils.py, tests/layers/test_layer_utils.py, model ValueErrors)."""
import numpy as np
import pytest

from handyrec_b200.features import DenseFeature, EmbdFeatureGroup, FeatureGroup, FeaturePool, SparseFeature, SparseSeqFeature
from handyrec_b200.features.utils import split_features
from handyrec_b200.keras_lite import Input
from handyrec_b200.layers import CustomEmbedding, SequencePoolingLayer
from handyrec_b200.layers.utils import concat, get_activation_layer, sampledsoftmaxloss

DEEPFM_CFG = {"FeatureGroups": {
    "fm_feature_group": {"type": "FeatureGroup", "name": "FM", "DenseFeatures": {"year": {"dim": 1, "dtype": "int32"}},
                         "SparseFeatures": {k: {"embedding_dim": 8} for k in ("user_id", "gender", "occupation", "zip", "age", "movie_id")},
                         "SparseSeqFeatures": {"hist_movie": {"unit": {"movie_id": {"embedding_dim": 8}}, "seq_len": 2},
                                               "genres": {"unit": {"genre_id": {"embedding_dim": 8}}, "seq_len": 3}}},
    "dnn_feature_group": {"type": "FeatureGroup", "name": "DNN", "DenseFeatures": {"year": {"dim": 1, "dtype": "int32"}},
                          "SparseFeatures": {k: {"embedding_dim": 8} for k in ("user_id", "gender", "occupation", "zip", "age", "movie_id")},
                          "SparseSeqFeatures": {"hist_movie": {"unit": {"movie_id": {"embedding_dim": 8}}, "seq_len": 2},
                                                "genres": {"unit": {"genre_id": {"embedding_dim": 8}}, "seq_len": 3}}}}}
DIN_CFG = {"FeatureGroups": {
    "item_seq_feat_group": {"type": "FeatureGroup", "name": "item_seq",
                            "SparseSeqFeatures": {"hist_movie": {"unit": {"movie_id": {"embedding_dim": 8}}, "seq_len": 2}}},
    "other_feature_group": {"type": "FeatureGroup", "name": "other_feats",
                            "SparseFeatures": {k: {"embedding_dim": 8} for k in ("user_id", "gender", "occupation", "zip", "age", "movie_id", "year")},
                            "SparseSeqFeatures": {"genres": {"unit": {"genre_id": {"embedding_dim": 8}}, "seq_len": 3}}}}}
FEATURE_DIM = {"user_id": 6, "gender": 3, "occupation": 21, "zip": 6, "age": 57, "movie_id": 11, "genre_id": 19, "year": 82}


def test_split_features_func():
    dense_feats = [DenseFeature("d1", dim=1), DenseFeature("d2", dim=2)]
    sparse_feats = [SparseFeature("s1", vocab_size=10, embedding_dim=16), SparseFeature("s2", vocab_size=20, embedding_dim=16)]
    sparse_seq_feats = [SparseSeqFeature(sparse_feats[0], "s1_seq", seq_len=15)]
    d, s, q = split_features(dense_feats + sparse_feats + sparse_seq_feats)
    assert d["d1"] == dense_feats[0] and d["d2"] == dense_feats[1]
    assert s["s1"] == sparse_feats[0] and s["s2"] == sparse_feats[1] and q["s1_seq"] == sparse_seq_feats[0]
    assert sparse_feats[0].embdding_dim == 16 and sparse_seq_feats[0].dtype == "int32" and not sparse_seq_feats[0].is_group


def test_FeatureGroup():
    dense_feats = [DenseFeature("d1", dim=1), DenseFeature("d2", dim=2)]
    sparse_feats = [SparseFeature("s1", vocab_size=10, embedding_dim=16), SparseFeature("s2", vocab_size=20, embedding_dim=16)]
    sparse_seq_feats = [SparseSeqFeature(SparseFeature("s3", vocab_size=24, embedding_dim=16), "s1_seq", seq_len=15)]
    pool = FeaturePool({"s1": np.random.rand(10, 16)})
    fg = FeatureGroup("FG", dense_feats + sparse_feats + sparse_seq_feats, pool)
    dense_outputs, embd_outputs = fg.embedding_lookup()
    assert dense_outputs[0].shape[1:] == (1,) and dense_outputs[1].shape[1:] == (2,)
    for embd_out in embd_outputs:
        assert embd_out.shape[1:] == (1, 16)
    # the unit of a sequence feature gets a mask_zero table, plain features do not (group.py:272-292)
    assert fg.embd_layers["s3"].mask_zero and not fg.embd_layers["s1"].mask_zero
    assert np.allclose(fg.embd_layers["s1"].get_weights()[0], pool.pre_embd["s1"].astype(np.float32))


def test_EmbdFeatureGroup():
    dense_feats = [DenseFeature("d1", dim=1), DenseFeature("d2", dim=2)]
    sparse_feats = [SparseFeature("s1", vocab_size=10, embedding_dim=16), SparseFeature("s2", vocab_size=20, embedding_dim=16)]
    sparse_seq_feats = [SparseSeqFeature(sparse_feats[0], "s2_seq", seq_len=4)]
    value_dict = {"s1": [1, 2, 3], "s2": [11, 12, 13], "d1": [-1, -2, -3], "d2": [[1, 2], [3, 4], [5, 6]],
                  "s2_seq": [[4, 5, 6, 7], [5, 6, 7, 8], [6, 7, 8, 9]]}
    pool = FeaturePool()
    with pytest.raises(ValueError):
        EmbdFeatureGroup("FG", "XXX", dense_feats + sparse_feats + sparse_seq_feats, pool, value_dict, embd_dim=8)
    with pytest.raises(ValueError):
        bad = [SparseSeqFeature(dense_feats[0], "d1_seq", seq_len=4)]
        EmbdFeatureGroup("FG", "s1", dense_feats + sparse_feats + bad, pool, value_dict, embd_dim=8)
    fg = EmbdFeatureGroup("FG", "s1", dense_feats + sparse_feats + sparse_seq_feats, pool, value_dict, embd_dim=8)
    item_id = fg.id_input
    assert fg.get_embd(item_id, compress=False).shape == (3, 3 * 16 + 1 + 2)
    assert fg.get_embd(item_id, compress=True).shape == (3, 8)
    seq_id_input = Input(shape=(3,), dtype="int32")
    assert fg.lookup(seq_id_input, compress=False).shape == (None, 3, 3 * 16 + 1 + 2)
    lo2 = fg.lookup(item_id, compress=False)
    assert lo2.shape[0] is None and lo2.shape[1] == 3 * 16 + 1 + 2
    call_outputs, call_mask = fg(seq_id_input)
    assert call_outputs.shape == (None, 3, 8) and call_mask.shape == (None, 3, 8)


def test_FeaturePool_init_input():
    pool = FeaturePool()
    pool.init_input("d1", {"name": "d1", "shape": (1,), "dtype": "float32"})
    with pytest.raises(AttributeError):
        pool.init_input("d1", {"name": "d1", "shape": (2,), "dtype": "float32"})
    with pytest.raises(AttributeError):
        pool.init_input("d1", {"name": "d1", "shape": (1,), "dtype": "int32"})


def test_FeaturePool_init_embd():
    pool = FeaturePool()
    params = {"name": "embd_s1", "input_dim": 12, "output_dim": 128, "trainable": True, "weights": None, "mask_zero": False}
    first = pool.init_embd("s1", params)
    params["mask_zero"] = True
    second = pool.init_embd("s1", params)
    assert second is not first and second.mask_zero  # H7: the mask-zero upgrade creates a NEW table (group.py:113-121)
    with pytest.raises(AttributeError):
        params["input_dim"] = 24
        pool.init_embd("s1", params)
    with pytest.raises(AttributeError):
        params["output_dim"] = 64
        pool.init_embd("s1", params)


def test_FeaturePool_init_pool():
    pool = FeaturePool()
    pool.init_pool("POOL_1", {"name": "POOL_1", "method": "mean"})
    with pytest.raises(AttributeError):
        pool.init_pool("POOL_1", {"name": "POOL_1", "method": "max"})
    with pytest.raises(AssertionError):
        SequencePoolingLayer("median")


def test_concat_func():
    dense_inputs = [Input(shape=(1,), dtype="int32"), Input(shape=(1,), dtype="float32")]
    embd_inputs = [CustomEmbedding(10, 64)(Input(shape=(1,), dtype="int32")) for _ in range(3)]
    with pytest.raises(ValueError):
        concat([], [])
    output = concat(dense_inputs, embd_inputs, keepdims=False)
    assert output.shape[0] is None and output.shape[1] == 2 * 1 + 64 * 3
    output = concat([], embd_inputs, axis=1, keepdims=True)
    assert output.shape[0] is None and output.shape[1:] == (3, 64)
    output = concat([dense_inputs[0]], [])
    assert output.shape[0] is None and output.shape[1] == 1
    output = concat([], embd_inputs)
    assert output.shape[0] is None and output.shape[1] == 64 * 3


def test_sampledsoftmaxloss_func():
    assert sampledsoftmaxloss([], [1, 2, 3]) == 2  # reference tests/layers/test_layer_utils.py:39


def test_get_activation_layer_func():
    get_activation_layer("dice")
    get_activation_layer("sigmoid")


def test_models_construct_and_raise():
    from handyrec_b200.config import ConfigLoader
    from handyrec_b200.models import DIN, DeepFM

    g = ConfigLoader(DEEPFM_CFG).prepare_features(FEATURE_DIM)
    with pytest.raises(ValueError):  # reference tests/models/ranking/context_aware/test_DeepFM.py:28-37
        DeepFM(g["fm_feature_group"], g["dnn_feature_group"], dnn_hidden_units=(8, 4), dnn_dropout=0.2, l2_dnn=0.2, dnn_bn=True)
    with pytest.warns(UserWarning):
        m = DeepFM(g["fm_feature_group"], g["dnn_feature_group"], dnn_hidden_units=(8, 1), dnn_dropout=0.2, l2_dnn=0.2, dnn_bn=True)
    assert [i.name for i in m.inputs] == ["year", "user_id", "gender", "occupation", "zip", "age", "movie_id", "hist_movie", "genres"]
    assert m.outputs.shape == (None, 1)
    # both groups share every table through the pool (DeepFM looks each one up twice in the reference)
    assert g["fm_feature_group"].embd_layers["movie_id"] is g["dnn_feature_group"].embd_layers["movie_id"]
    g2 = ConfigLoader(DIN_CFG).prepare_features(FEATURE_DIM)
    m2 = DIN(g2["item_seq_feat_group"], g2["other_feature_group"], dnn_hidden_units=(8,), lau_dnn_hidden_units=(8, 1))
    assert m2.outputs.shape == (None, 1) and "hist_movie" in [i.name for i in m2.inputs]
    # H7: item_seq group comes first with mask_zero=True; the other group reuses that masked table
    assert g2["other_feature_group"].embd_layers["movie_id"] is g2["item_seq_feat_group"].embd_layers["movie_id"]


# ------------------------------------------------------------------------------------------------
# lowering: which graphs the fused engine takes (handyrec_b200/lowering.py) -- pattern matching only, no kernel runs
# ------------------------------------------------------------------------------------------------
def _deepfm(fm_feats, dnn_feats, pool=None, **kw):
    from handyrec_b200.models import DeepFM

    pool = pool or FeaturePool()
    return DeepFM(FeatureGroup("fm", fm_feats, pool, l2_embd=0.0), FeatureGroup("dnn", dnn_feats, pool, l2_embd=0.0), **kw)


def test_lowering_accepts_the_deepfm_graphs_the_engine_covers():
    from handyrec_b200.lowering import match_deepfm

    sparse = [SparseFeature("a", 50, 8), SparseFeature("b", 7, 8)]
    hist = SparseSeqFeature(SparseFeature("item", 30, 8), "hist", 5)
    dense = [DenseFeature("d0"), DenseFeature("d1")]
    spec = match_deepfm(_deepfm(sparse + [hist], dense + sparse + [hist], dnn_hidden_units=(16, 8, 1)))
    assert spec is not None
    assert [(f.inp.name, f.seq_len, f.pool) for f in spec.fields] == [("a", 1, "none"), ("b", 1, "none"), ("hist", 5, "mean")]
    assert [t.name for t in spec.dense_inputs] == ["d0", "d1"] and tuple(spec.dnn.hidden_units) == (16, 8, 1)
    # no dense features at all; sigmoid / tanh towers
    assert match_deepfm(_deepfm(sparse, sparse, dnn_hidden_units=(8, 1), dnn_activation="tanh")) is not None
    # a table shared by a plain and a sequence feature is ONE engine table with two fields
    item = SparseFeature("item", 30, 8)
    shared = match_deepfm(_deepfm([item, SparseSeqFeature(item, "hist", 4)], [item, SparseSeqFeature(item, "hist", 4)], dnn_hidden_units=(8, 1)))
    assert shared is not None and shared.fields[0].emb is shared.fields[1].emb


@pytest.mark.parametrize("case", ["bn", "dropout", "dice", "different_groups", "mixed_dims", "dim_not_multiple_of_4"])
def test_lowering_declines_what_the_engine_does_not_cover(case):
    """Declined graphs keep the layer-by-layer path (same results, more launches): BatchNorm / Dropout in the tower, Dice, FM and DNN
    groups over different features (the engine shares ONE lookup), more than one embedding dim, a dim the 16-byte row chunks do not divide."""
    from handyrec_b200.lowering import match_deepfm

    a, b, c = SparseFeature("a", 50, 8), SparseFeature("b", 7, 8), SparseFeature("c", 9, 8)
    if case == "bn":
        m = _deepfm([a, b], [a, b], dnn_hidden_units=(8, 1), dnn_bn=True)
    elif case == "dropout":
        m = _deepfm([a, b], [a, b], dnn_hidden_units=(8, 1), dnn_dropout=0.3)
    elif case == "dice":
        m = _deepfm([a, b], [a, b], dnn_hidden_units=(8, 1), dnn_activation="dice")
    elif case == "different_groups":
        m = _deepfm([a, b], [a, b, c], dnn_hidden_units=(8, 1))
    elif case == "mixed_dims":
        try:
            m = _deepfm([a, SparseFeature("w", 5, 16)], [a, SparseFeature("w", 5, 16)], dnn_hidden_units=(8, 1))
        except Exception:  # the FM layer itself needs one embedding dim (interaction.py:26-39): nothing to lower
            m = None
    else:
        m = _deepfm([SparseFeature("p", 5, 6), SparseFeature("q", 5, 6)], [SparseFeature("p", 5, 6), SparseFeature("q", 5, 6)], dnn_hidden_units=(8, 1))
    if m is None:
        return  # such a DeepFM cannot even be built
    assert match_deepfm(m) is None


def test_lowering_declines_other_model_families():
    from handyrec_b200.config import ConfigLoader
    from handyrec_b200.lowering import match_deepfm
    from handyrec_b200.models import DIN

    g = ConfigLoader(DIN_CFG).prepare_features(FEATURE_DIM)
    m = DIN(g["item_seq_feat_group"], g["other_feature_group"], dnn_hidden_units=(8,), lau_dnn_hidden_units=(8, 1))
    assert match_deepfm(m) is None
