"""GPU: N emulated ranks (threads, one GPU, no inter-kernel waiting) run the CUDA row exchange and the sharded engine."""
import threading

import pytest
import torch

import oracle
from tests import shard_helpers as H
from tests.test_gpu_parity import close

pytestmark = pytest.mark.gpu


def _run_threads(n, fn):
    errs, outs = [], [None] * n

    def wrap(r):
        try:
            torch.cuda.set_device(0)
            outs[r] = fn(r)
        except Exception as e:  # noqa: BLE001
            import traceback

            errs.append((r, traceback.format_exc()))
            raise

    ts = [threading.Thread(target=wrap, args=(r,)) for r in range(n)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not errs, errs[0][1]
    return outs


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("opt", ["sgd"])
def test_row_exchange_cuda(dev, world, opt):
    from handyrec_b200 import kernels as K
    from handyrec_b200._lib import OptParams
    from handyrec_b200.sharded import CudaShardProvider, RowExchange

    B = 65
    vocabs, tables, fields, ids, douts, D = H.make_case(world, B, seed=5)
    shared = H.ThreadComm.Shared(world)
    shards = [[t.to(dev) for t in H.shards_of(tables, r, world)] for r in range(world)]

    def rank_fn(r):
        plan = K.LookupPlan(shards[r], fields)
        ex = RowExchange(CudaShardProvider(plan, vocabs, world, B), H.ThreadComm(shared, r))
        out = torch.zeros(B, 5 * D, device=dev)
        ex.forward(ids[r].to(dev), out)
        op = OptParams()
        op.opt, op.lr = 0, 0.5
        ex.backward_update(ids[r].to(dev), douts[r].to(dev), op)
        torch.cuda.synchronize()
        return out.cpu()

    outs = _run_threads(world, rank_fn)
    want_tabs = H.expected_tables_after_sgd(tables, fields, ids, douts, 0.5)
    for r in range(world):
        want = H.expected_forward(tables, fields, ids[r])
        assert torch.equal(outs[r][:, : 3 * D], want[:, : 3 * D])
        close(outs[r], want, 1e-5)
        for s, w in zip(shards[r], want_tabs):
            ws = w[r::world]
            if ws.shape[0]:
                close(s[: ws.shape[0]], ws, 1e-4)


@pytest.mark.parametrize("world,b,peer,rep,opt", [(2, 64, True, 0, "sgd"), (4, 64, False, 0, "sgd"), (2, 512, True, 0, "sgd"), (2, 512, False, 0, "sgd"),
                                                  (8, 64, True, 0, "sgd"), (2, 512, True, 60, "adam"), (4, 64, False, 60, "adam"), (2, 64, True, 60, "sgd"),
                                                  (2, 512, True, 1000, "adam")])
def test_sharded_engine_matches_single_gpu_engine(dev, world, b, peer, rep, opt):
    """N emulated ranks (large tables row-sharded, tables <= rep rows replicated) take the same step as one engine on the
    concatenated batch."""
    from handyrec_b200 import kernels as K
    from handyrec_b200.engine import DeepFMEngine
    from handyrec_b200.sharded import ShardedDeepFMEngine

    D, n_dense, hidden = 8, 3, (16, 1)  # b = 512 takes the tcgen05 path with the overlapped embedding backward
    g = torch.Generator().manual_seed(1)
    vocabs = [11, 50, 300]
    tables = [(torch.rand(v, D, generator=g) - 0.5) * 0.1 for v in vocabs]
    fields = [(0, 1, "none"), (1, 1, "none"), (2, 1, "none")]
    ids = [torch.stack([torch.randint(0, v, (b,), generator=g, dtype=torch.int32) for v in vocabs], 1) for _ in range(world)]
    dense = [torch.randn(b, n_dense, generator=g) for _ in range(world)]
    label = [(torch.rand(b, generator=g) < 0.3).float() for _ in range(world)]
    ref = DeepFMEngine([t.clone().to(dev) for t in tables], fields, n_dense, hidden, "relu", batch_size=b * world, optimizer=opt, lr=0.1, seed=7,
                       dense_table_max_rows=rep)
    ref.train_step_on_device(torch.cat(ids).to(dev), torch.cat(dense).to(dev), torch.cat(label).to(dev))
    torch.cuda.synchronize()
    shared = H.ThreadComm.Shared(world)
    engines = [None] * world

    def rank_fn(r):
        sh = [(full if v <= rep else s_).to(dev) for full, s_, v in zip(tables, H.shards_of(tables, r, world), vocabs)]
        comm = H.ThreadComm(shared, r)
        peer_ptrs = comm.share_ptrs(sh) if peer else None  # peer-mapped forward (no all-to-all) vs. row-exchange forward
        eng = ShardedDeepFMEngine(sh, vocabs, fields, n_dense, comm, peer_ptrs=peer_ptrs, dnn_hidden_units=hidden, dnn_activation="relu",
                                  batch_size=b, optimizer=opt, lr=0.1, seed=7, replicate_max_rows=rep)
        assert eng.peer_lookup == peer or rep >= 1000
        engines[r] = eng
        eng.train_step_on_device(ids[r].to(dev), dense[r].to(dev), label[r].to(dev))
        torch.cuda.synchronize()
        return float(eng.loss_sum)

    losses = _run_threads(world, rank_fn)
    assert abs(sum(losses) - float(ref.loss_sum)) < 1e-4 * abs(float(ref.loss_sum))
    for r in range(world):
        close(engines[r].params, ref.params, 1e-4)  # replicated dense parameters stay identical after the all-reduce
        for t in range(len(vocabs)):
            ws = ref.tables[t] if vocabs[t] <= rep else ref.tables[t][r::world]
            if ws.shape[0]:
                close(engines[r].tables[t][: ws.shape[0]], ws, 1e-4)
