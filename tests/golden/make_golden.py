"""Regenerates tests/golden/c1_deepfm.npz from the CPU oracle (python tests/golden/make_golden.py).

The reference (TF 2.6) cannot be imported in this image, so these vectors pin the ORACLE, not the reference: they
freeze the restatement's numbers for the ml-1m-test DeepFM shape (B=5, 6 sparse + hist_movie L=2 + genres L=3, D=8,
one int32 dense feature; /root/reference/tests/ml-1m-test/DeepFM_cfg.yaml) so that any later change of the oracle or
of the CUDA path shows up against a committed file."""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import oracle  # noqa: E402

VOCABS = OrderedDict(user_id=6, gender=3, occupation=21, zip=6, age=57, movie_id=11, genre_id=19)


def build():
    D, B = 8, 5
    tables = OrderedDict((n, oracle.hash_uniform_table(v, D, seed=100 + i)) for i, (n, v) in enumerate(VOCABS.items()))
    g = np.random.RandomState(2022)
    x = OrderedDict((n, g.randint(0, VOCABS[n], size=(B, 1)).astype(np.int32)) for n in ("user_id", "gender", "occupation", "zip", "age", "movie_id"))
    x["year"] = g.randint(0, 80, size=(B, 1)).astype(np.int32)
    x["hist_movie"] = np.array([[0, 0], [0, 3], [1, 2], [0, 10], [7, 7]], dtype=np.int32)
    x["genres"] = np.array([[1, 5, 0], [2, 0, 0], [18, 3, 4], [0, 0, 0], [6, 6, 0]], dtype=np.int32)
    p = oracle.dnn_init(1 + 8 * D, (8, 1), seed=7)
    fm_w = torch.from_numpy(oracle.hash_uniform_table(D, 1, seed=55, lo=-0.5, hi=0.5))
    fm_w0 = torch.tensor([0.125])
    T = {n: torch.from_numpy(t) for n, t in tables.items()}
    sparse = OrderedDict((n, (T[n], torch.from_numpy(x[n]), n == "movie_id")) for n in ("user_id", "gender", "occupation", "zip", "age", "movie_id"))
    seqs = OrderedDict((("hist_movie", (T["movie_id"], torch.from_numpy(x["hist_movie"]))), ("genres", (T["genre_id"], torch.from_numpy(x["genres"])))))
    out = {}
    for pool in ("mean", "sum", "max"):
        out[f"pooled_{pool}"] = oracle.concat([], oracle.group_embedding_lookup(sparse, seqs, pool)).numpy()
    embds = oracle.group_embedding_lookup(sparse, seqs, "mean")
    fm_in = oracle.concat([], embds, axis=1, keepdims=True)
    out["fm"] = oracle.fm(fm_in, fm_w, fm_w0).numpy()
    out["prob"] = oracle.deepfm_forward([torch.from_numpy(x["year"])], embds, embds, p, fm_w, fm_w0).numpy()
    arrays = {f"x_{k}": v for k, v in x.items()}
    arrays.update({f"table_{k}": v for k, v in tables.items()})
    arrays.update({f"dnn_W{i}": w.numpy() for i, w in enumerate(p.W)})
    arrays.update({f"dnn_b{i}": b.numpy() for i, b in enumerate(p.b)})
    arrays.update(fm_w=fm_w.numpy(), fm_w0=fm_w0.numpy(), **out)
    return arrays


def build_din():
    """DIN local-activation attention at the ml-1m-test DIN shape in miniature (layers/sequence.py:92-102 + tools.py:104-113 +
    models/ranking/sequential/DIN.py:87-93): B=6 users, history length T=5 (one user with an all-padding history, one full),
    D=8, LAU tower (16, 8, 1) with Dice and with relu."""
    D, B, T, V = 8, 6, 5, 40
    table = torch.from_numpy(oracle.hash_uniform_table(V, D, seed=300, lo=-0.5, hi=0.5))
    qid = np.array([[3], [7], [1], [39], [12], [5]], dtype=np.int32)
    kid = np.array([[4, 9, 0, 0, 0], [0, 0, 0, 0, 0], [1, 2, 3, 4, 5], [39, 39, 0, 0, 0], [8, 0, 0, 0, 0], [5, 6, 7, 0, 0]], dtype=np.int32)
    out = {"din_table": table.numpy(), "din_qid": qid, "din_kid": kid}
    for act in ("dice", "relu"):
        p = oracle.dnn_init(4 * D, (16, 8, 1), seed=11)
        for i, u in enumerate(p.units):
            p.b[i] = torch.from_numpy(oracle.hash_uniform_table(1, u, seed=400 + i, lo=-0.1, hi=0.1))[0]
            p.dice_alpha[i] = torch.from_numpy(oracle.hash_uniform_table(1, u, seed=410 + i, lo=-0.3, hi=0.3))[0]
            p.dice_mean[i] = torch.from_numpy(oracle.hash_uniform_table(1, u, seed=420 + i, lo=-0.1, hi=0.1))[0]
            p.dice_var[i] = torch.from_numpy(oracle.hash_uniform_table(1, u, seed=430 + i, lo=0.5, hi=1.5))[0]
        q, _ = oracle.custom_embedding(table, torch.from_numpy(qid), True)
        keys, kmask = oracle.custom_embedding(table, torch.from_numpy(kid), True)
        keys, kmask = oracle.squeeze_mask(keys, kmask)
        score = oracle.local_activation_unit(q, keys, kmask, p, act=act, training=False)
        out[f"din_score_{act}"] = score.numpy()
        out[f"din_pooled_{act}"] = oracle.din_attention_pool(score, keys).numpy()
        if act == "dice":
            for i in range(len(p.units)):
                out[f"din_W{i}"], out[f"din_b{i}"] = p.W[i].numpy(), p.b[i].numpy()
                out[f"din_alpha{i}"], out[f"din_mean{i}"], out[f"din_var{i}"] = p.dice_alpha[i].numpy(), p.dice_mean[i].numpy(), p.dice_var[i].numpy()
    return out


if __name__ == "__main__":
    here = os.path.dirname(__file__)
    np.savez_compressed(os.path.join(here, "c1_deepfm.npz"), **build())
    np.savez_compressed(os.path.join(here, "c2_din.npz"), **build_din())
    print("wrote c1_deepfm.npz, c2_din.npz")
