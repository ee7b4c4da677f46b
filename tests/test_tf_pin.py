"""Pins the oracle to the REAL reference the day TensorFlow is importable.

TensorFlow is not in this image, so today every test below is skipped and DESIGN.md says "parity unpinned".
The moment `import tensorflow` works next to a checkout of the reference (REFERENCE_ROOT, default /root/reference), these tests
build the genuine `handyrec.layers.{CustomEmbedding, SequencePoolingLayer, FM, LocalActivationUnit, DNN, Dice}` with fixed
weights, run them on the CPU and compare outputs and gradients with `oracle/` on the same inputs -- the check BASELINE.md
promises.  CPU-only (no `gpu` marker): it validates the checker, not the CUDA path.
"""
import os
import sys

import numpy as np
import pytest

tf = pytest.importorskip("tensorflow", reason="TensorFlow is not installed in this image: parity stays unpinned (DESIGN.md 4)")
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REF, "handyrec")):
    pytest.skip("no reference checkout next to TensorFlow", allow_module_level=True)
if REF not in sys.path:
    sys.path.insert(0, REF)

import torch  # noqa: E402

import oracle  # noqa: E402

FWD_TOL, GRAD_TOL = 1e-5, 1e-4


def _close(got, want, tol):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=tol, atol=tol * max(float(np.abs(want).max()), 1e-30))


def _rng(seed):
    return np.random.RandomState(seed)


def test_custom_embedding_and_pooling_match_reference():
    from handyrec.layers import CustomEmbedding, SequencePoolingLayer

    r = _rng(0)
    V, D, B, L = 23, 8, 17, 5
    W = r.uniform(-0.05, 0.05, (V, D)).astype(np.float32)
    ids = r.randint(0, V, (B, L)).astype(np.int32)
    ids[r.rand(B, L) < 0.4] = 0
    ids[3] = 0
    emb = CustomEmbedding(V, D, weights=[W], mask_zero=True)
    x = emb(tf.constant(ids))
    mask = emb.compute_mask(tf.constant(ids))
    o_x, o_mask = oracle.custom_embedding(torch.from_numpy(W), torch.from_numpy(ids), True)
    assert np.array_equal(x.numpy(), o_x.numpy())            # gathers are bit-exact
    assert np.array_equal(mask.numpy(), o_mask.numpy())      # tiled != 0 mask
    for method in ("mean", "sum", "max"):
        got = SequencePoolingLayer(method)(x, mask=mask)
        _close(got.numpy(), oracle.sequence_pooling(o_x, o_mask, method).numpy(), FWD_TOL)


def test_fm_matches_reference():
    from handyrec.layers import FM

    r = _rng(1)
    B, F, D = 9, 6, 8
    x = r.randn(B, F, D).astype(np.float32)
    layer = FM()
    xt = tf.constant(x)
    with tf.GradientTape() as tape:
        tape.watch(xt)
        y = layer(xt)
        loss = tf.reduce_sum(y * y)
    dx = tape.gradient(loss, xt)
    w, w0 = None, None
    for v in layer.weights:
        if int(np.prod(v.shape)) == D:
            w = v.numpy().reshape(D, 1)
        elif int(np.prod(v.shape)) == 1:
            w0 = v.numpy().reshape(1)
    assert w is not None and w0 is not None
    xo = torch.from_numpy(x).requires_grad_(True)
    yo = oracle.fm(xo, torch.from_numpy(w), torch.from_numpy(w0))
    (yo * yo).sum().backward()
    _close(y.numpy(), yo.detach().numpy(), FWD_TOL)
    _close(dx.numpy(), xo.grad.numpy(), GRAD_TOL)


def _dnn_params_from_keras(dnn_layer):
    """(W_i, b_i) of the Dense layers and (alpha, mean, var) of the Dice layers of a built reference DNN, in call order."""
    Ws, bs, dice = [], [], []
    for l in dnn_layer.layers.layers if hasattr(dnn_layer.layers, "layers") else dnn_layer.layers:
        name = type(l).__name__
        if name == "Dense":
            Ws.append(l.kernel.numpy())
            bs.append(l.bias.numpy())
        elif name == "Dice":
            dice.append((l.alpha.numpy(), l.bn.moving_mean.numpy(), l.bn.moving_variance.numpy()))
    return Ws, bs, dice


@pytest.mark.parametrize("act", ["relu", "sigmoid", "dice"])
def test_dnn_and_dice_match_reference(act):
    from handyrec.layers import DNN

    r = _rng(2)
    x = r.randn(33, 12).astype(np.float32)
    layer = DNN(hidden_units=(16, 8, 1), activation=act, output_activation="linear")
    y = layer(tf.constant(x), training=False)
    Ws, bs, dice = _dnn_params_from_keras(layer)
    p = oracle.dnn_init(12, (16, 8, 1))
    p.W = [torch.from_numpy(w) for w in Ws]
    p.b = [torch.from_numpy(b) for b in bs]
    for i, d in enumerate(dice):  # hidden layers carry a Dice each (the last layer has the linear output activation)
        p.dice_alpha[i], p.dice_mean[i], p.dice_var[i] = (torch.from_numpy(t) for t in d)
    yo = oracle.dnn(torch.from_numpy(x), p, act=act, output_activation="linear", training=False)
    _close(y.numpy(), yo.detach().numpy(), FWD_TOL)


def test_local_activation_unit_matches_reference():
    from handyrec.layers import LocalActivationUnit

    r = _rng(3)
    B, T, D = 7, 5, 4
    q = r.randn(B, 1, D).astype(np.float32)
    k = r.randn(B, T, D).astype(np.float32)
    m = r.rand(B, T) > 0.3
    lau = LocalActivationUnit(hidden_units=(8, 1), activation="sigmoid")
    y = lau([tf.constant(q), tf.constant(k)], mask=[None, tf.constant(m)])
    Ws, bs, _ = _dnn_params_from_keras(lau.dnn)
    p = oracle.dnn_init(4 * D, (8, 1))
    p.W = [torch.from_numpy(w) for w in Ws]
    p.b = [torch.from_numpy(b) for b in bs]
    yo = oracle.local_activation_unit(torch.from_numpy(q), torch.from_numpy(k), torch.from_numpy(m), p, act="sigmoid")
    _close(y.numpy(), yo.detach().numpy(), FWD_TOL)
