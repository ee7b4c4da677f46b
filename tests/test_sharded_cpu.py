"""CPU (gloo, world size 2): the row-exchange protocol of handyrec_b200.sharded against the single-rank oracle."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import shard_helpers as H


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from handyrec_b200._lib import OptParams
        from handyrec_b200.sharded import RowExchange, TorchDistComm

        vocabs, tables, fields, ids, douts, D = H.make_case(world, B, seed=3)
        shards = H.shards_of(tables, rank, world)
        ex = RowExchange(H.CpuShardProvider(shards, vocabs, fields, world, rank), TorchDistComm())
        out = torch.zeros(B, 5 * D)
        ex.forward(ids[rank], out)
        want = H.expected_forward(tables, fields, ids[rank])
        fwd_err = float((out - want).abs().max())
        exact = bool(torch.equal(out[:, : 3 * D], want[:, : 3 * D]))  # plain gathers survive the exchange bit-exactly
        op = OptParams()
        op.opt, op.lr = 0, 0.5
        ex.backward_update(ids[rank], douts[rank], op)
        want_tabs = H.expected_tables_after_sgd(tables, fields, ids, douts, 0.5)
        upd_err = max(float((s - w[rank::world][: s.shape[0]]).abs().max()) for s, w in zip(shards, want_tabs) if w[rank::world].shape[0])
        q.put((rank, fwd_err, exact, upd_err, ex.n_send, ex.n_recv))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_row_exchange_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 17, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(r[4] for r in res) == sum(r[5] for r in res)  # everything sent is received
    for rank, fwd_err, exact, upd_err, n_send, n_recv in res:
        assert exact, rank
        assert fwd_err < 1e-6, (rank, fwd_err)
        assert upd_err < 1e-5, (rank, upd_err)


def test_key_base_and_shard_rows():
    from handyrec_b200.sharded import key_base_table, shard_rows

    assert [shard_rows(10, r, 4) for r in range(4)] == [3, 3, 2, 2]
    assert [shard_rows(3, r, 8) for r in range(8)] == [1, 1, 1, 1, 1, 1, 1, 1]  # empty shards still hold one (unused) row
    kb = key_base_table([10, 3], 4)
    assert kb[:, 0].tolist() == [0, 0, 0, 0] and kb[:, 1].tolist() == [3, 3, 2, 2]


def test_exchange_offsets_reproduce_the_all_to_all_layout():
    """The fused gather + exchange writes rank s's rows for owner d at exchange_offsets(...)[2][d] of d's receive buffer.  With every
    rank doing that independently, each owner must hold exactly what an all-to-all ordered by sender would have delivered."""
    import numpy as np

    from handyrec_b200.sharded import exchange_offsets

    rng = np.random.RandomState(0)
    for n in (1, 2, 3, 8):
        C = rng.randint(0, 50, (n, n)).tolist()
        if n > 1:
            C[1][0] = 0  # an empty segment
        payload = {(s, d): [(s, d, i) for i in range(C[s][d])] for s in range(n) for d in range(n)}
        buffers = [[None] * sum(C[s][d] for s in range(n)) for d in range(n)]
        for s in range(n):
            send, recv, off = exchange_offsets(C, s)
            assert send == C[s] and recv == [C[r][s] for r in range(n)]
            for d in range(n):
                for i, item in enumerate(payload[(s, d)]):
                    assert buffers[d][off[d] + i] is None  # no two senders write the same row
                    buffers[d][off[d] + i] = item
        for d in range(n):
            assert buffers[d] == [item for s in range(n) for item in payload[(s, d)]]
