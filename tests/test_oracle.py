"""CPU: the oracle against its known-answer vectors and internal consistency (no GPU)."""
import numpy as np
import torch

import oracle
from oracle import kat


def test_kats():
    kat.self_test()


def test_hash_table_deterministic_and_in_range():
    a = oracle.hash_uniform_table(100, 16, seed=7)
    b = oracle.hash_uniform_table(100, 16, seed=7)
    c = oracle.hash_uniform_table(100, 16, seed=8)
    assert a.dtype == np.float32 and np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.min() >= -0.05 and a.max() < 0.05 and abs(a.mean()) < 5e-3
    # a row % N shard is a strided view of the full table
    s = oracle.hash_uniform_table(25, 16, seed=7, row_start=3, row_step=4)
    assert np.array_equal(s, a[3::4])


def test_sharded_lookup_matches_single_rank():
    g = torch.Generator().manual_seed(0)
    table = torch.randn(97, 8, generator=g)
    ids = torch.randint(0, 97, (33, 7), generator=g, dtype=torch.int32)
    ids[ids % 3 == 0] = 0
    ids[5] = 0
    for method in ("mean", "sum", "max"):
        seq, mask = oracle.custom_embedding(table, ids, True)
        want = oracle.sequence_pooling(seq, mask, method)
        for n in (1, 2, 4, 8):
            got = oracle.sharded_lookup_emulated(table, ids, n, method)
            if method == "max":
                # all-padding rows are ~-1e9 in both
                np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-3)
            else:
                np.testing.assert_allclose(got.numpy(), want.numpy(), rtol=1e-5, atol=1e-6)


def test_embedding_grad_formula_matches_autograd():
    g = torch.Generator().manual_seed(1)
    W = torch.randn(11, 4, generator=g, requires_grad=True)
    ids = torch.randint(0, 11, (6, 3), generator=g)
    out, mask = oracle.custom_embedding(W, ids, True)
    pooled = oracle.sequence_pooling(out, mask, "sum")
    l2 = 1e-3
    loss = (pooled * pooled).sum() + l2 * (W * W).sum()
    loss.backward()
    dout = (2 * pooled.detach()).expand(-1, 3, -1) * mask.float()
    want = oracle.embedding_grad_dense(11, ids, dout, W.detach(), l2)
    np.testing.assert_allclose(W.grad.numpy(), want.numpy(), rtol=1e-5, atol=1e-6)


def test_dnn_structure_quirks():
    p = oracle.dnn_init(5, (4, 1), seed=0)
    assert [w.shape for w in p.W] == [(5, 5), (5, 4), (4, 1)]  # extra Dense(in) (core.py:57)
    x = torch.randn(3, 5)
    y = oracle.dnn(x, p, act="relu", output_activation="linear")
    ref = torch.relu(torch.relu(x @ p.W[0] + p.b[0]) @ p.W[1] + p.b[1]) @ p.W[2] + p.b[2]
    np.testing.assert_allclose(y.numpy(), ref.numpy(), rtol=1e-6)
    # activation falsy -> the elif applies output_activation on every layer (core.py:66-69)
    y2 = oracle.dnn(x, p, act=None, output_activation="sigmoid")
    ref2 = torch.sigmoid(torch.sigmoid(torch.sigmoid(x @ p.W[0] + p.b[0]) @ p.W[1] + p.b[1]) @ p.W[2] + p.b[2])
    np.testing.assert_allclose(y2.numpy(), ref2.numpy(), rtol=1e-6)


def test_concat_layout_and_errors():
    import pytest

    with pytest.raises(ValueError):
        oracle.concat([], [])
    d = [torch.ones(2, 1, dtype=torch.int32), torch.ones(2, 1)]
    e = [torch.zeros(2, 1, 4) for _ in range(3)]
    out = oracle.concat(d, e)
    assert out.shape == (2, 2 + 12) and out.dtype == torch.float32 and out[0, 0] == 1 and out[0, 2] == 0  # dense first
    assert oracle.concat([], e, axis=1, keepdims=True).shape == (2, 3, 4)
    assert oracle.concat([d[0]], []).shape == (2, 1)
    assert oracle.concat([], e).shape == (2, 12)


def test_log_uniform_sampler_distribution_and_uniqueness():
    """Host sampler of SampledSoftmaxLayer vs the closed-form log-uniform probabilities; draws are unique; the oracle's restatement
    and the layer's vectorised sampler follow the same law."""
    import numpy as np

    from handyrec_b200.layers.tools import LogUniformSampler

    C = 50
    s = LogUniformSampler(C, seed=1)
    counts = np.zeros(C)
    n = 20000
    for _ in range(n):
        ids, tries = s.sample(1)
        assert tries >= 1
        counts[ids[0]] += 1
    p = oracle.log_uniform_prob(np.arange(C), C)
    assert abs(p.sum() - 1.0) < 1e-12
    assert np.abs(counts / n - p).max() < 4 * np.sqrt(p.max() / n)
    ids, tries = s.sample(20)
    assert len(set(ids.tolist())) == 20 and tries >= 20 and ids.dtype == np.int32
    ids, _ = LogUniformSampler(7, seed=0).sample(100)  # more candidates than classes: every class once
    assert sorted(ids.tolist()) == list(range(7))
    e = oracle.unique_expected_count(p, 30)
    assert np.all((e > 0) & (e <= 1)) and np.all(np.diff(e) < 0)


def test_long_run_classification_is_a_local_property():
    """The sorted backward takes "long" rows out of its chunk/merge chain.  The detect kernel lists a row when its run covers two
    consecutive sample points; the chunk kernel decides per position from the four sample points around it.  Both must agree for
    every position, whatever the run layout: skewed, runs ending exactly on sample points, sentinel tails, n not a multiple of the step."""
    from oracle.layers_ref import long_run_keys_global, long_run_local_test

    rng = np.random.RandomState(5)
    step, sentinel = 8, 10 ** 6
    for trial in range(300):
        n = int(rng.randint(1, 200))
        kind = trial % 4
        if kind == 0:
            keys = rng.randint(0, 6, n)
        elif kind == 1:  # a few very long runs between singletons
            keys = np.repeat(rng.randint(0, 50, 12), rng.randint(1, 40, 12))[:n]
        elif kind == 2:  # run boundaries exactly on sample points
            keys = np.repeat(np.arange(40), step * rng.randint(1, 4))[:n]
        else:
            keys = np.where(rng.rand(n) < 0.2, sentinel, rng.randint(0, 3, n))
        keys = np.sort(keys)
        long_rows = long_run_keys_global(keys, step, sentinel)
        for i in range(len(keys)):
            assert long_run_local_test(keys, step, sentinel, i) == (int(keys[i]) in long_rows), (trial, i, keys.tolist())
        # every listed row is listed once by the detect rule "the first pair of the run reports it"
        pts = np.arange(0, len(keys), step)
        first_pairs = [int(keys[pts[j]]) for j in range(len(pts) - 1)
                       if keys[pts[j]] == keys[pts[j + 1]] and keys[pts[j]] != sentinel and (j == 0 or keys[pts[j - 1]] != keys[pts[j]])]
        assert sorted(first_pairs) == sorted(long_rows)
        assert len(long_rows) <= len(keys) // (2 * step) + 2  # the bound the workspace is sized with
