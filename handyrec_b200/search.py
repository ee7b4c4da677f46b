"""Top-n inner-product search and ranking metrics.

`search_embedding` has the signature of handyrec/models/utils.py:7-51 (which wraps faiss.IndexFlatIP): exact top-n by inner
product, ties broken by the lower item index.  Items stay on the device; their scores are computed chunk by chunk with the
library's GEMM (queries . items^T: the tcgen05 3xTF32 kernel from 512 queries on) and folded into per-query running lists by
`hrb_topk_merge`.  `map_at_k / recall_at_k / hr_at_k` restate handyrec/data/metrics.py on the host.
"""
from __future__ import annotations

from typing import Any, Tuple

import numpy as np
import torch

from . import _lib
from . import kernels as K
from ._lib import call

ITEM_CHUNK = 8192


def topn_inner_product(queries: torch.Tensor, items: torch.Tensor, n: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """queries (Q, D), items (N, D) fp32 CUDA tensors -> (scores (Q, n), indices (Q, n) int32), best first."""
    Q, D = queries.shape
    N = items.shape[0]
    n = min(int(n), N)
    queries, items = queries.contiguous(), items.contiguous()
    val = torch.empty(Q, n, device=queries.device, dtype=torch.float32)
    idx = torch.empty(Q, n, device=queries.device, dtype=torch.int32)
    call("hrb_topk_init", K._p(val), K._p(idx), Q, n, K._stream())
    use_tc = Q >= 512 and D >= 16 and D % 4 == 0
    buf = torch.empty(Q, min(ITEM_CHUNK, N), device=queries.device, dtype=torch.float32)
    for c0 in range(0, N, ITEM_CHUNK):
        c1 = min(N, c0 + ITEM_CHUNK)
        chunk = items[c0:c1]
        out = buf[:, : c1 - c0]
        if use_tc and c1 - c0 >= 16 and (c1 - c0) % 4 == 0:
            K.dense_bwd_x_t(queries, chunk, out=out)               # queries (Q,D) . chunk (C,D)^T on the tensor cores
        else:
            K.dense_bwd_x(queries, chunk, out=out, mode=_lib.GEMM_FP32)
        call("hrb_topk_merge", K._p(out), buf.stride(0), Q, c1 - c0, c0, n, K._p(val), K._p(idx), K._stream())
    return val, idx


def search_embedding(embd_dim: int, item_embd: np.ndarray, user_embd: np.ndarray, item_list: np.ndarray, n: int, gpu: bool = True) -> np.ndarray:
    """handyrec/models/utils.py:7-51: top-n items for each user embedding -> (NUM_USERS, n) array of entries of `item_list`."""
    dev = torch.device("cuda", torch.cuda.current_device())
    items = torch.as_tensor(np.ascontiguousarray(item_embd, dtype=np.float32)).to(dev)
    users = torch.as_tensor(np.ascontiguousarray(user_embd, dtype=np.float32)).to(dev)
    if items.shape[1] != embd_dim or users.shape[1] != embd_dim:
        raise ValueError("embedding dimension mismatch")
    _, idx = topn_inner_product(users, items, n)
    return np.asarray(item_list)[idx.cpu().numpy().astype(np.int64)]


# ---- handyrec/data/metrics.py -------------------------------------------------------------------
def _apk(actual, predicted, k=10):
    predicted = list(predicted)[:k]
    score, hits = 0.0, 0.0
    for i, p in enumerate(predicted):
        if p in actual and p not in predicted[:i]:
            hits += 1.0
            score += hits / (i + 1.0)
    return score / min(len(actual), k)


def map_at_k(actual: Any, predicted: Any, k: int = 12) -> float:
    return float(np.mean([_apk(a, p, k) for a, p in zip(actual, predicted)]))


def recall_at_k(actual: Any, predicted: Any, k: int = 12) -> float:
    return float(np.mean([sum(1 for r in a if r in list(p)[:k]) / len(a) for a, p in zip(actual, predicted)]))


def hr_at_k(actual: Any, predicted: Any, k: int = 10) -> float:
    return sum(1 for a, p in zip(actual, predicted) if any(x in a for x in list(p)[:k])) / len(actual)
