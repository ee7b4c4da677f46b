"""Known-answer vectors for the hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference's own layer tests are empty files
(/root/reference/tests/layers/test_{core,interaction,sequence,tools,activation}.py
are 0 lines) and TensorFlow cannot be imported here, so these vectors were
derived BY HAND from the formulas in the reference sources:

* KAT-pool   handyrec/layers/tools.py:93-101 + handyrec/layers/sequence.py:26-46
* KAT-FM     handyrec/layers/interaction.py:26-39
* KAT-LAU    handyrec/layers/sequence.py:92-102 + models/ranking/sequential/DIN.py:93
* KAT-Dice   handyrec/layers/activation.py:27-42
* KAT-ssl    tests/layers/test_layer_utils.py:39  (the one numeric assertion the
             reference's tests hold: ``sampledsoftmaxloss([], [1,2,3]) == 2``)

``self_test()`` checks the oracle against every vector; tests/test_oracle.py
runs it, and the GPU parity tests run the same vectors through the C ABI.
"""
from __future__ import annotations

import numpy as np
import torch

from . import layers_ref as R

F32 = np.float32

# ---- KAT-pool -------------------------------------------------------------
POOL_TABLE = np.array([[2 * r - 2, 2 * r - 1.5, 2 * r - 1, 2 * r - 0.5] for r in range(5)], dtype=F32)
POOL_IDS = np.array([[3, 0, 1], [0, 0, 0], [4, 4, 2]], dtype=np.int32)
POOL_MASK = POOL_IDS != 0
POOL_MEAN = np.array([[2, 2.5, 3, 3.5], [0, 0, 0, 0], [4.6666665, 5.166667, 5.666667, 6.166667]], dtype=F32)
POOL_SUM = np.array([[4, 5, 6, 7], [0, 0, 0, 0], [14, 15.5, 17, 18.5]], dtype=F32)
POOL_MAX = np.array([[4, 4.5, 5, 5.5], [-1e9] * 4, [6, 6.5, 7, 7.5]], dtype=F32)

# ---- KAT-FM ---------------------------------------------------------------
FM_X = np.array([[[1, 2], [3, 4], [5, 6]], [[0.5, -1], [0, 0], [2, 1]]], dtype=F32)
FM_W = np.array([[0.1], [-0.2]], dtype=F32)
FM_W0 = np.array([0.25], dtype=F32)
FM_PART2 = np.array([-1.5, 0.25], dtype=F32)
FM_PART3 = np.array([67.0, 0.0], dtype=F32)
FM_OUT = np.array([[65.75], [0.5]], dtype=F32)
FM_DX = np.array(
    [[[8.1, 9.8], [6.1, 7.8], [4.1, 5.8]], [[2.1, 0.8], [2.6, -0.2], [0.6, -1.2]]], dtype=F32
)  # d(sum out)/dX[b,f,:] = w^T + (sum_f X[b] - X[b,f])

# ---- KAT-LAU input --------------------------------------------------------
LAU_Q = np.array([[[1, 2]]], dtype=F32)  # (1,1,2)
LAU_K = np.array([[[0, 0], [1, 1], [2, -1]]], dtype=F32)  # (1,3,2)
LAU_MASK = np.array([[False, True, True]])
LAU_ATT_IN = np.array(
    [[[1, 2, 0, 0, 1, 2, 0, 0], [1, 2, 1, 1, 0, 1, 1, 2], [1, 2, 2, -1, -1, 3, 2, -2]]], dtype=F32
)
LAU_ATT = np.array([[[0.0, 0.9, 0.6]]], dtype=F32)  # stand-in score 0.1*sum(att_in), masked
LAU_POOLED = np.array([[[2.1, 0.3]]], dtype=F32)

# ---- KAT-Dice (inference; mean 0, var 1, eps 1e-9, alpha .25) ---------------
DICE_X = np.array([[-2.0, -0.5, 0.0, 1.5]], dtype=F32)
DICE_OUT = np.array([[-0.6788044, -0.26657775, 0.0, 1.2947713]], dtype=F32)


def lau_standin_params() -> R.DNNParams:
    """A DNN whose net effect is ``0.1 * sum(att_in)``: identity Dense(8) then Dense(1) of 0.1s (act=None)."""
    p = R.DNNParams(8, (1,))
    p.W = [torch.eye(8), torch.full((8, 1), 0.1)]
    p.b = [torch.zeros(8), torch.zeros(1)]
    for lst in (p.dice_alpha, p.dice_mean, p.dice_var, p.bn_gamma, p.bn_beta, p.bn_mean, p.bn_var):
        lst.extend([None, None])
    return p


def self_test() -> None:
    T = R.as_t
    table = T(POOL_TABLE)
    seq, mask = R.custom_embedding(table, POOL_IDS, mask_zero=True)
    assert seq.shape == (3, 3, 4) and mask.shape == (3, 3, 4) and mask.dtype == torch.bool
    assert np.array_equal(seq.numpy(), POOL_TABLE[POOL_IDS])  # bit-exact gather
    assert np.array_equal(mask.numpy(), np.repeat(POOL_MASK[..., None], 4, -1))  # bit-exact mask
    for method, want in (("mean", POOL_MEAN), ("sum", POOL_SUM), ("max", POOL_MAX)):
        got = R.sequence_pooling(seq, mask, method)
        assert got.shape == (3, 1, 4)
        np.testing.assert_allclose(got[:, 0].numpy(), want, rtol=1e-6, atol=0)
    _, nomask = R.custom_embedding(table, POOL_IDS, mask_zero=False)
    assert nomask is None
    try:
        R.sequence_pooling(seq, None, "mean")
        raise AssertionError("expected ValueError")
    except ValueError:
        pass

    x = T(FM_X).clone().requires_grad_(True)
    out = R.fm(x, T(FM_W), T(FM_W0))
    np.testing.assert_allclose(out.detach().numpy(), FM_OUT, rtol=1e-6)
    out.sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), FM_DX, rtol=1e-6, atol=1e-6)

    queries = T(LAU_Q).repeat_interleave(3, dim=1)
    att_in = torch.cat([queries, T(LAU_K), queries - T(LAU_K), queries * T(LAU_K)], -1)
    assert np.array_equal(att_in.numpy(), LAU_ATT_IN)
    att = R.local_activation_unit(T(LAU_Q), T(LAU_K), T(LAU_MASK), lau_standin_params(), act=None)
    np.testing.assert_allclose(att.numpy(), LAU_ATT, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(R.din_attention_pool(att, T(LAU_K)).numpy(), LAU_POOLED, rtol=1e-6, atol=1e-7)

    d = R.dice(T(DICE_X), torch.full((4,), 0.25), torch.zeros(4), torch.ones(4), training=False)
    np.testing.assert_allclose(d.numpy(), DICE_OUT, rtol=1e-6, atol=1e-7)

    assert float(R.sampledsoftmaxloss([], [1, 2, 3])) == 2.0  # reference tests/layers/test_layer_utils.py:39


if __name__ == "__main__":
    self_test()
    print("oracle KATs OK")
