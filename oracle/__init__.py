"""CPU oracle for the HandyRec hot path.  TEST INFRASTRUCTURE ONLY.

This package is an op-for-op CPU restatement (torch-CPU fp32 + a little numpy)
of the reference's lookup -> pool -> FM / LAU / DNN path.  It exists to CHECK
the CUDA product in ``handyrec_b200`` and to be timed as the CPU baseline:

* only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
  ``cpu_baseline`` / ``--impl reference`` legs may import it;
* nothing under ``handyrec_b200/`` imports it, and the product never falls back
  to it (the product raises when its CUDA library is missing).

PARITY UNPINNED: the reference (TF 2.6 / Keras) cannot be imported in this
image (no tensorflow wheel, no network) and its own layer tests are empty files,
so there is no golden vector from the reference itself.  The oracle is pinned by
(i) the formulas in the reference sources cited on every function
(``/root/reference/handyrec/...`` file:line) plus TF's documented op semantics,
(ii) the hand-derived known-answer vectors in ``oracle/kat.py``
and (iii) the only numeric assertion the reference tests hold near the path
(``tests/layers/test_layer_utils.py:39``).
"""
from .layers_ref import *  # noqa: F401,F403
