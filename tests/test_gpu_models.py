"""GPU: the Keras-like mirror end to end -- DeepFM / DIN built from the reference's YAML schema, forward and one
optimiser step against the oracle with identical weights; fit/predict smoke runs shaped like the reference's model tests."""
from collections import OrderedDict

import numpy as np
import pytest
import torch

import oracle
from tests.test_gpu_parity import close, rand_ids
from tests.test_mirror_cpu import DEEPFM_CFG, DIN_CFG, FEATURE_DIM

pytestmark = pytest.mark.gpu


def _data(B, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = {k: torch.randint(0, FEATURE_DIM[k], (B, 1), generator=g, dtype=torch.int32).numpy() for k in ("user_id", "gender", "occupation", "zip", "age", "movie_id")}
    x["year"] = torch.randint(0, FEATURE_DIM["year"], (B, 1), generator=g, dtype=torch.int32).numpy()
    x["hist_movie"] = rand_ids(B, 2, FEATURE_DIM["movie_id"], seed=seed + 1).numpy()
    x["genres"] = rand_ids(B, 3, FEATURE_DIM["genre_id"], seed=seed + 2, pre_pad=False).numpy()
    y = (torch.rand(B, generator=g) < 0.4).float().numpy()
    return x, y


def _dnn_params(dnn_layer, in_dim, hidden):
    from handyrec_b200.layers.core import Dense

    p = oracle.DNNParams(in_dim, hidden)
    for l in dnn_layer.layers:
        if isinstance(l, Dense):
            p.W.append(l.kernel.detach().cpu().clone().requires_grad_(True))
            p.b.append(l.bias.detach().cpu().clone().requires_grad_(True))
    from handyrec_b200.layers.activation import Dice

    dices = [l for l in dnn_layer.layers if isinstance(l, Dice)]
    for i in range(len(p.units)):
        d = dices[i] if i < len(dices) else None
        p.dice_alpha.append(d.alphas.detach().cpu().clone().requires_grad_(True) if d else None)
        p.dice_mean.append(d.moving_mean.detach().cpu().clone() if d else None)
        p.dice_var.append(d.moving_variance.detach().cpu().clone() if d else None)
    return p


def _oracle_deepfm(groups, model, x, tables):
    sparse_names = ["user_id", "gender", "occupation", "zip", "age", "movie_id"]
    sparse = OrderedDict((n, (tables[n], torch.from_numpy(x[n]), n == "movie_id")) for n in sparse_names)
    seqs = OrderedDict((("hist_movie", (tables["movie_id"], torch.from_numpy(x["hist_movie"]))), ("genres", (tables["genre_id"], torch.from_numpy(x["genres"])))))
    embds = oracle.group_embedding_lookup(sparse, seqs, "mean")
    embds2 = oracle.group_embedding_lookup(sparse, seqs, "mean")  # the DNN group looks everything up again (DeepFM.py:63)
    deep = model.get_layer("Deep_Part")
    p = _dnn_params(deep, 1 + 8 * 8, (8, 1))
    fm = model.get_layer("FM_Part")
    fm_w = fm.linear.detach().cpu().clone().requires_grad_(True)
    fm_w0 = fm.w_0.detach().cpu().clone().requires_grad_(True)
    logit = oracle.deepfm_forward([torch.from_numpy(x["year"])], embds, embds2, p, fm_w, fm_w0, return_logit=True)
    return logit, p, fm_w, fm_w0


def test_deepfm_model_forward_and_sgd_step(dev):
    from handyrec_b200.config import ConfigLoader
    from handyrec_b200.keras_lite import SGD, binary_crossentropy
    from handyrec_b200.models import DeepFM

    g = ConfigLoader(DEEPFM_CFG).prepare_features(FEATURE_DIM)
    with pytest.warns(UserWarning):
        model = DeepFM(g["fm_feature_group"], g["dnn_feature_group"], dnn_hidden_units=(8, 1))
    x, y = _data(64)
    pool = g["feature_pool"]
    prob = model.predict(x)  # builds the weights
    tables = {n: l.embeddings.detach().cpu().clone().requires_grad_(True) for n, l in pool.embd_layers.items()}
    logit, p, fm_w, fm_w0 = _oracle_deepfm(g, model, x, tables)
    close(prob, torch.sigmoid(logit).detach(), 1e-5)
    assert prob.shape == (64, 1)
    # one SGD step, l2_embd = 1e-6 (group.py:189) on every table
    l2 = 1e-6
    loss = oracle.bce_from_logits(logit[:, 0], torch.from_numpy(y)) + sum(l2 * (t * t).sum() for t in tables.values())
    loss.backward()
    model.compile(optimizer=SGD(lr=0.5), loss=binary_crossentropy)
    got = model.train_on_batch(x, y)
    assert abs(got - float(loss.detach())) < 1e-5 * max(1.0, abs(float(loss.detach())))
    for n, l in pool.embd_layers.items():
        close(l.embeddings.detach(), tables[n].detach() - 0.5 * tables[n].grad, 1e-4)
    from handyrec_b200.layers.core import Dense

    denses = [l for l in model.get_layer("Deep_Part").layers if isinstance(l, Dense)]
    for i, l in enumerate(denses):
        close(l.kernel.detach(), p.W[i].detach() - 0.5 * p.W[i].grad, 1e-4)
        close(l.bias.detach(), p.b[i].detach() - 0.5 * p.b[i].grad, 1e-4)
    fm = model.get_layer("FM_Part")
    close(fm.linear.detach(), fm_w.detach() - 0.5 * fm_w.grad, 1e-4)
    close(fm.w_0.detach(), fm_w0.detach() - 0.5 * fm_w0.grad, 1e-4)


def test_deepfm_fit_like_reference_test(dev):
    """tests/models/ranking/context_aware/test_DeepFM.py:39-54: bn + dropout + l2, Adam(1e-4), 2 epochs at batch 5, predict."""
    from handyrec_b200.config import ConfigLoader
    from handyrec_b200.keras_lite import Adam, binary_crossentropy
    from handyrec_b200.models import DeepFM

    g = ConfigLoader(DEEPFM_CFG).prepare_features(FEATURE_DIM)
    with pytest.warns(UserWarning):
        model = DeepFM(g["fm_feature_group"], g["dnn_feature_group"], dnn_hidden_units=(8, 1), dnn_dropout=0.2, l2_dnn=0.2, dnn_bn=True)
    x, y = _data(23)
    model.compile(optimizer=Adam(lr=1e-4), loss=binary_crossentropy)
    hist = model.fit(x=x, y=y, batch_size=5, epochs=2, validation_data=(x, y))
    assert len(hist.history["loss"]) == 2 and all(np.isfinite(hist.history["loss"])) and len(hist.history["val_loss"]) == 2
    pred = model.predict(x, batch_size=5)
    assert pred.shape == (23, 1) and np.all((pred > 0) & (pred < 1))


@pytest.mark.parametrize("act", ["dice", "sigmoid"])
def test_din_model_forward_and_gradients(dev, act):
    from handyrec_b200.config import ConfigLoader
    from handyrec_b200.keras_lite import SGD, binary_crossentropy
    from handyrec_b200.layers import LocalActivationUnit
    from handyrec_b200.models import DIN

    g = ConfigLoader(DIN_CFG).prepare_features(FEATURE_DIM)
    model = DIN(g["item_seq_feat_group"], g["other_feature_group"], dnn_hidden_units=(8,), dnn_activation=act, lau_dnn_hidden_units=(8, 1),
                lau_dnn_activation=act)
    x, y = _data(48, seed=3)
    x["hist_movie"][0] = 0  # an all-padding history
    prob = model.predict(x)
    pool = g["feature_pool"]
    # non-trivial Dice parameters so the test sees them
    lau = [l for l in model.layers if isinstance(l, LocalActivationUnit)][0]
    from handyrec_b200.layers.activation import Dice

    tgen = torch.Generator().manual_seed(5)
    for l in model.layers:
        for s in [l] + list(getattr(l, "_sublayers", [])) + [ss for s2 in getattr(l, "_sublayers", []) for ss in getattr(s2, "_sublayers", [])]:
            if isinstance(s, Dice):
                s.alphas.data.copy_((torch.rand(s.alphas.shape, generator=tgen) * 0.5).to(dev))
    prob = model.predict(x)
    tables = {n: l.embeddings.detach().cpu().clone().requires_grad_(True) for n, l in pool.embd_layers.items()}

    def oracle_forward(training):
        names = ["user_id", "gender", "occupation", "zip", "age", "movie_id", "year"]
        sparse = OrderedDict((n, (tables[n], torch.from_numpy(x[n]), n == "movie_id")) for n in names)
        seqs = OrderedDict((("genres", (tables["genre_id"], torch.from_numpy(x["genres"]))),))
        other = oracle.group_embedding_lookup(sparse, seqs, "mean")
        keys, kmask = oracle.custom_embedding(tables["movie_id"], torch.from_numpy(x["hist_movie"]), True)
        keys, kmask = oracle.squeeze_mask(keys, kmask)
        q, _ = oracle.custom_embedding(tables["movie_id"], torch.from_numpy(x["movie_id"]), True)
        p_lau = _dnn_params(lau.dnn, 32, (8, 1))
        att = oracle.local_activation_unit(q, keys, kmask, p_lau, act=act, training=training)
        pooled = oracle.din_attention_pool(att, keys)
        dnn_layer = [l for l in model.layers if type(l).__name__ == "DNN"][0]
        p = _dnn_params(dnn_layer, 9 * 8, (8, 1))
        dnn_in = oracle.concat([], other + [pooled])
        logit = oracle.dnn(dnn_in, p, act=act, output_activation=None, training=training)
        return logit, p, p_lau

    logit, _, _ = oracle_forward(False)
    close(prob, torch.sigmoid(logit).detach(), 1e-5)
    # training step (Dice uses batch statistics), SGD, gradients of every table / kernel vs autograd of the oracle
    logit, p, p_lau = oracle_forward(True)
    loss = oracle.bce_from_logits(logit[:, 0], torch.from_numpy(y)) + sum(1e-6 * (t * t).sum() for t in tables.values())
    loss.backward()
    model.compile(optimizer=SGD(lr=1.0), loss=binary_crossentropy)
    got = model.train_on_batch(x, y)
    assert abs(got - float(loss.detach())) < 1e-5 * max(1.0, abs(float(loss.detach())))
    for n, l in pool.embd_layers.items():
        close(l.embeddings.detach(), tables[n].detach() - tables[n].grad, 1e-4)
    from handyrec_b200.layers.core import Dense

    for lay, pp in ((lau.dnn, p_lau), ([l for l in model.layers if type(l).__name__ == "DNN"][0], p)):
        denses = [l for l in lay.layers if isinstance(l, Dense)]
        for i, l in enumerate(denses):
            close(l.kernel.detach(), pp.W[i].detach() - pp.W[i].grad, 1e-4)
        dices = [l for l in lay.layers if isinstance(l, Dice)]
        for i, l in enumerate(dices):
            close(l.alphas.detach(), pp.dice_alpha[i].detach() - pp.dice_alpha[i].grad, 1e-4)


def test_layer_face_eager_calls(dev):
    """Layers called directly on CUDA tensors (no graph), like Keras layers in eager mode."""
    from handyrec_b200.layers import FM, CustomEmbedding, SequencePoolingLayer

    emb = CustomEmbedding(50, 8, mask_zero=True)
    ids = rand_ids(33, 5, 50, seed=1).to(dev)
    seq = emb(ids)
    assert seq.shape == (33, 5, 8) and seq._keras_mask.shape == (33, 5, 8)
    table = emb.embeddings.detach().cpu()
    want_seq, want_mask = oracle.custom_embedding(table, ids.cpu(), True)
    assert torch.equal(seq.detach().cpu(), want_seq) and torch.equal(seq._keras_mask.cpu(), want_mask)
    pooled = SequencePoolingLayer("mean")(seq)
    close(pooled, oracle.sequence_pooling(want_seq, want_mask, "mean"), 1e-5)
    nomask = CustomEmbedding(50, 8)(ids)
    with pytest.raises(ValueError):  # sequence.py:27-28
        SequencePoolingLayer("sum")(nomask)
    fm = FM()
    xin = torch.randn(33, 6, 8, device=dev)
    out = fm(xin)
    close(out, oracle.fm(xin.cpu(), fm.linear.detach().cpu(), fm.w_0.detach().cpu()), 1e-5)


def test_embd_feature_group_values(dev):
    """a12: EmbdFeatureGroup.get_embd / lookup / __call__ (features/group.py:439-516) against the oracle."""
    from handyrec_b200.features import DenseFeature, EmbdFeatureGroup, FeaturePool, SparseFeature, SparseSeqFeature
    from handyrec_b200.keras_lite import Input, Model

    dense_feats = [DenseFeature("d1", dim=1), DenseFeature("d2", dim=2)]
    sparse_feats = [SparseFeature("s1", vocab_size=10, embedding_dim=16), SparseFeature("s2", vocab_size=20, embedding_dim=16)]
    seq_feats = [SparseSeqFeature(sparse_feats[0], "s2_seq", seq_len=4)]
    value_dict = {"s1": [0, 1, 2, 3], "s2": [11, 12, 13, 14], "d1": [-1.0, -2.0, -3.0, -4.0], "d2": [[1.0, 2], [3, 4], [5, 6], [7, 8]],
                  "s2_seq": [[0, 0, 0, 0], [5, 6, 0, 0], [6, 7, 8, 9], [0, 0, 0, 3]]}
    fg = EmbdFeatureGroup("FG", "s1", dense_feats + sparse_feats + seq_feats, FeaturePool(), value_dict, embd_dim=8, pool_method="mean")
    item_id = fg.id_input
    seq_in = Input(shape=(3,), name="seq", dtype="int32")
    full = fg.get_embd(item_id, compress=False)
    looked = fg.lookup(seq_in, compress=False)
    out, mask = fg(seq_in)
    model = Model(inputs=[item_id, seq_in], outputs=[full, looked, out, mask])
    ids = np.array([[1], [3]], dtype=np.int32)
    seq = np.array([[1, 2, 0], [3, 0, 1]], dtype=np.int32)
    got_full, got_looked, got_out, got_mask = model({"s1": ids, "seq": seq})
    t1 = fg.embd_layers["s1"].embeddings.detach().cpu()
    t2 = fg.embd_layers["s2"].embeddings.detach().cpu()
    e_s1 = t1[torch.tensor(value_dict["s1"])]
    e_s2 = t2[torch.tensor(value_dict["s2"])]
    seq_e, seq_m = oracle.custom_embedding(t1, torch.tensor(value_dict["s2_seq"]), True)
    e_seq = oracle.sequence_pooling(seq_e, seq_m, "mean")[:, 0]
    want_full = torch.cat([torch.tensor(value_dict["d1"]).unsqueeze(-1), torch.tensor(value_dict["d2"]), e_s1, e_s2, e_seq], -1)
    assert got_full.shape == (4, 1 + 2 + 48)
    close(got_full, want_full, 1e-5)
    close(got_looked, want_full[torch.from_numpy(seq).long()], 1e-5)
    w, b = fg._output_layer.kernel.detach().cpu(), fg._output_layer.bias.detach().cpu()
    close(got_out, (want_full @ w + b)[torch.from_numpy(seq).long()], 1e-5)
    assert got_mask.shape == (2, 3, 8) and np.array_equal(got_mask.cpu().numpy(), np.repeat((seq != 0)[..., None], 8, -1))
