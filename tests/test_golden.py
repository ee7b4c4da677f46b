"""Golden fixture tests/golden/c1_deepfm.npz (made by tests/golden/make_golden.py from the oracle; parity with the
reference itself is unpinned, see oracle/__init__.py): the oracle still reproduces it (CPU), and the CUDA path matches it (GPU)."""
import os
import sys
from collections import OrderedDict

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
GOLD = os.path.join(HERE, "golden", "c1_deepfm.npz")
SPARSE = ("user_id", "gender", "occupation", "zip", "age", "movie_id")


def test_oracle_reproduces_golden():
    import make_golden

    gold = np.load(GOLD)
    now = make_golden.build()
    assert set(now) == set(gold.files)
    for k in gold.files:
        np.testing.assert_array_equal(now[k], gold[k], err_msg=k)
    # hand-checkable entries: all-padding rows pool to 0 (mean/sum) and ~-1e9 (max) (layers/sequence.py:35-46)
    assert np.all(gold["pooled_mean"][0, 48:56] == 0) and np.all(gold["pooled_mean"][3, 56:64] == 0)
    assert np.allclose(gold["pooled_max"][3, 56:64], -1e9)


@pytest.mark.gpu
def test_cuda_matches_golden(dev):
    from handyrec_b200 import kernels as K
    from handyrec_b200.engine import DeepFMEngine

    gold = np.load(GOLD)
    names = list(SPARSE) + ["genre_id"]
    tables = [torch.from_numpy(gold[f"table_{n}"]).to(dev) for n in names]
    ids = torch.from_numpy(np.concatenate([gold[f"x_{n}"] for n in SPARSE] + [gold["x_hist_movie"], gold["x_genres"]], 1)).to(dev)
    D = 8
    for pool in ("mean", "sum", "max"):
        fields = [(i, 1, "none", i, i * D) for i in range(6)] + [(5, 2, pool, 6, 6 * D), (6, 3, pool, 8, 7 * D)]
        got = K.LookupPlan(tables, fields).forward(ids)["out"].cpu().numpy()
        want = gold[f"pooled_{pool}"]
        assert np.array_equal(got[:, : 6 * D], want[:, : 6 * D])  # plain gathers: bit-exact
        np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-7)
    eng = DeepFMEngine(tables, [(i, 1, "none") for i in range(6)] + [(5, 2, "mean"), (6, 3, "mean")], 1, (8, 1), "relu", batch_size=5, optimizer="sgd")
    for i in range(3):
        eng.set_dense_weights(i, torch.from_numpy(gold[f"dnn_W{i}"]), torch.from_numpy(gold[f"dnn_b{i}"]))
    eng.fm_w.copy_(torch.from_numpy(gold["fm_w"][:, 0]).to(dev))
    eng.fm_w0.copy_(torch.from_numpy(gold["fm_w0"]).to(dev))
    prob = eng.predict_on_device(ids, torch.from_numpy(gold["x_year"]).to(dev)).cpu().numpy()
    np.testing.assert_allclose(prob, gold["prob"][:, 0], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(eng.fm_out.cpu().numpy(), gold["fm"][:, 0], rtol=1e-5, atol=1e-6)


GOLD_DIN = os.path.join(HERE, "golden", "c2_din.npz")


def test_oracle_reproduces_din_golden():
    import make_golden

    gold = np.load(GOLD_DIN)
    now = make_golden.build_din()
    assert set(now) == set(gold.files)
    for k in gold.files:
        np.testing.assert_array_equal(now[k], gold[k], err_msg=k)
    # hand-checkable: padded positions score exactly 0 (att_out * mask, layers/sequence.py:101); the user with an all-padding
    # history pools to the zero vector (tf.matmul(att, keys), DIN.py:93); there is no softmax, so scores need not sum to 1
    for act in ("dice", "relu"):
        s, pooled = gold[f"din_score_{act}"], gold[f"din_pooled_{act}"]
        assert s.shape == (6, 1, 5) and pooled.shape == (6, 1, 8)
        assert np.all(s[:, 0][gold["din_kid"] == 0] == 0) and np.all(pooled[1] == 0)
        assert np.all(s[2, 0] != 0)
        # identical history items get identical scores (row 3: movie 39 twice)
        assert s[3, 0, 0] == s[3, 0, 1]


@pytest.mark.gpu
def test_cuda_matches_din_golden(dev):
    from handyrec_b200 import kernels as K

    gold = np.load(GOLD_DIN)
    table = torch.from_numpy(gold["din_table"]).to(dev)
    qid = torch.from_numpy(gold["din_qid"]).reshape(-1).to(dev)
    kid = torch.from_numpy(gold["din_kid"]).to(dev)
    W = [torch.from_numpy(gold[f"din_W{i}"]).to(dev) for i in range(4)]
    b = [torch.from_numpy(gold[f"din_b{i}"]).to(dev) for i in range(4)]
    dice = [tuple(torch.from_numpy(gold[f"din_{n}{i}"]).to(dev) for n in ("alpha", "mean", "var")) for i in range(4)]
    units = [w.shape[1] for w in W]
    for act in ("dice", "relu"):
        params = K.lau_pack_params(W, b, dice if act == "dice" else None)
        score, pooled = K.lau_fwd(table, qid, kid, params, units, act)
        got_s, got_p = score.cpu().numpy(), pooled.cpu().numpy()
        assert np.all(got_s[:, 0][gold["din_kid"] == 0] == 0)  # the mask is exact
        np.testing.assert_allclose(got_s, gold[f"din_score_{act}"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(got_p, gold[f"din_pooled_{act}"], rtol=1e-5, atol=1e-7)
