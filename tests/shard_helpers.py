"""Test infrastructure for the row-sharded lookup: an oracle-backed CPU provider (so the RowExchange protocol can run
under gloo without a GPU), a thread-based in-process communicator (N emulated ranks on one GPU) and the expected values."""
from __future__ import annotations

import threading
from collections import OrderedDict
from typing import List, Sequence

import torch

import oracle
from handyrec_b200.sharded import key_base_table, shard_rows


def make_case(n_ranks: int, B: int, seed: int = 0, D: int = 8):
    """Per-rank ids for a small group: 3 plain features + one mean-pooled and one sum-pooled sequence feature."""
    g = torch.Generator().manual_seed(seed)
    vocabs = [5, 37, 1000, 23]
    tables = [(torch.rand(v, D, generator=g) - 0.5) * 0.1 for v in vocabs]
    # (table, seq_len, pool, ids_col, out_col)
    fields = [(0, 1, "none", 0, 0), (1, 1, "none", 1, D), (2, 1, "none", 2, 2 * D), (3, 4, "mean", 3, 3 * D), (1, 3, "sum", 7, 4 * D)]
    ids = []
    for r in range(n_ranks):
        cols = [torch.randint(0, vocabs[t], (B, 1), generator=g, dtype=torch.int32) for t in (0, 1, 2)]
        h = torch.randint(1, vocabs[3], (B, 4), generator=g, dtype=torch.int32)
        h[torch.rand(B, 4, generator=g) < 0.4] = 0
        h[0] = 0
        s = torch.randint(1, vocabs[1], (B, 3), generator=g, dtype=torch.int32)
        s[torch.rand(B, 3, generator=g) < 0.3] = 0
        ids.append(torch.cat(cols + [h, s], 1))
    douts = [torch.randn(B, 5 * D, generator=g) for _ in range(n_ranks)]
    return vocabs, tables, fields, ids, douts, D


def shards_of(tables: Sequence[torch.Tensor], rank: int, n_ranks: int) -> List[torch.Tensor]:
    out = []
    for t in tables:
        s = t[rank::n_ranks].clone()
        if s.shape[0] == 0:
            s = torch.zeros(1, t.shape[1])
        out.append(s.contiguous())
    return out


def expected_forward(tables, fields, ids):
    sparse, seqs = OrderedDict(), OrderedDict()
    pools = {}
    for f, (t, L, pool, ic, oc) in enumerate(fields):
        if pool == "none":
            sparse[f"f{f}"] = (tables[t], ids[:, ic : ic + 1], False)
        else:
            seqs[f"f{f}"] = (tables[t], ids[:, ic : ic + L])
            pools[f"f{f}"] = pool
    outs = [oracle.custom_embedding(*v[:2], v[2])[0] for v in sparse.values()]
    for k, (tab, i) in seqs.items():
        outs.append(oracle.sequence_pooling(*oracle.custom_embedding(tab, i, True), pools[k]))
    return oracle.concat([], outs)


def expected_tables_after_sgd(tables, fields, ids_all, douts_all, lr):
    leaf = [t.clone().requires_grad_(True) for t in tables]
    loss = 0
    for ids, dout in zip(ids_all, douts_all):
        loss = loss + (expected_forward(leaf, fields, ids) * dout).sum()
    loss.backward()
    return [t - lr * l.grad for t, l in zip(tables, leaf)]


class CpuShardProvider:
    """Same contract as handyrec_b200.sharded.CudaShardProvider, in torch-CPU ops (TEST INFRASTRUCTURE)."""

    def __init__(self, shard_tables, vocabs, fields, n_ranks, rank):
        self.tables, self.vocabs, self.fields, self.N, self.rank = shard_tables, vocabs, fields, n_ranks, rank
        self.kb = key_base_table(vocabs, n_ranks)
        self.D = shard_tables[0].shape[1]
        self.P = sum(f[1] for f in fields)
        self.pos_field, self.pos_col = [], []
        for fi, f in enumerate(fields):
            self.pos_col.append(len(self.pos_field))
            self.pos_field += [fi] * f[1]

    def _pos(self, ids):
        B = ids.shape[0]
        idv = torch.stack([ids[:, self.fields[fi][3] + (c - self.pos_col[fi])] for c, fi in enumerate(self.pos_field)], 1).long()
        pool_none = torch.tensor([self.fields[fi][2] == "none" for fi in self.pos_field])
        table = torch.tensor([self.fields[fi][0] for fi in self.pos_field])
        return idv, pool_none.unsqueeze(0).expand(B, -1), table.unsqueeze(0).expand(B, -1)

    def route(self, ids):
        idv, pool_none, table = self._pos(ids)
        valid = pool_none | (idv != 0)
        owner = torch.where(valid, idv % self.N, torch.full_like(idv, self.N)).reshape(-1)
        key = (self.kb[owner.clamp(max=self.N - 1), table.reshape(-1)] + idv.reshape(-1) // self.N)
        perm = torch.argsort(owner, stable=True)
        counts = torch.bincount(owner, minlength=self.N + 1)
        return perm.to(torch.int32), key[perm].to(torch.int32), counts.to(torch.int64)

    def rows_by_key(self, keys, n):
        base = self.kb[self.rank]
        keys = keys[:n].long()
        t = torch.searchsorted(base, keys, right=True) - 1
        out = torch.zeros(n, self.D)
        for j in range(n):
            out[j] = self.tables[int(t[j])][int(keys[j] - base[t[j]])]
        return out

    def scatter_rows(self, ids, perm, n, rows, out):
        B = ids.shape[0]
        pos_rows = torch.zeros(B * self.P, self.D)
        for j in range(n):
            p = int(perm[j])
            b, c = divmod(p, self.P)
            f = self.fields[self.pos_field[c]]
            if f[2] == "none":
                out[b, f[4] : f[4] + self.D] = rows[j]
            else:
                pos_rows[p] = rows[j]
        for fi, f in enumerate(self.fields):
            if f[2] == "none":
                continue
            seq = pos_rows.view(B, self.P, self.D)[:, self.pos_col[fi] : self.pos_col[fi] + f[1]]
            mask = (ids[:, f[3] : f[3] + f[1]] != 0).unsqueeze(-1).expand(-1, -1, self.D)
            out[:, f[4] : f[4] + self.D] = oracle.sequence_pooling(seq, mask, f[2])[:, 0]

    def gather_grads(self, ids, perm, n, dout):
        send = torch.zeros(n, self.D)
        for j in range(n):
            b, c = divmod(int(perm[j]), self.P)
            f = self.fields[self.pos_field[c]]
            s = 1.0
            if f[2] == "mean":
                cnt = int((ids[b, f[3] : f[3] + f[1]] != 0).sum())
                s = 1.0 / cnt if cnt else 0.0
            send[j] = dout[b, f[4] : f[4] + self.D] * s
        return send

    def keyed_update(self, keys, grads, n, op):
        base = self.kb[self.rank]
        keys = keys[:n].long()
        t = torch.searchsorted(base, keys, right=True) - 1
        for ti, tab in enumerate(self.tables):
            sel = t == ti
            if sel.any():
                g = torch.zeros(tab.shape, dtype=torch.float64)
                g.index_add_(0, keys[sel] - base[ti], grads[:n][sel].double())
                tab -= (op.lr * g).float()


class ThreadComm:
    """In-process emulation of N ranks (one thread each) for single-GPU tests: data moves through shared slots."""

    class Shared:
        def __init__(self, n):
            self.n, self.barrier, self.slots = n, threading.Barrier(n), [[None] * n for _ in range(n)]

    def __init__(self, shared, rank):
        self.s, self.rank, self.N = shared, rank, shared.n

    def _exchange(self, pieces):
        if isinstance(pieces[0], torch.Tensor) and pieces[0].is_cuda:  # emulated ranks share one GPU but use different streams: publish only finished data
            torch.cuda.current_stream().synchronize()
        for dst, t in enumerate(pieces):
            self.s.slots[self.rank][dst] = t
        self.s.barrier.wait()
        got = [self.s.slots[src][self.rank] for src in range(self.N)]
        got = [g.clone() if isinstance(g, torch.Tensor) else g for g in got]
        self.s.barrier.wait()
        return got

    def all_to_all_equal(self, send):
        return torch.cat(self._exchange(list(send.chunk(self.N))))

    def all_to_all_v(self, send, send_counts, recv_counts):
        got = self._exchange(list(torch.split(send, send_counts)))
        assert [g.shape[0] for g in got] == recv_counts
        return torch.cat(got)

    def all_reduce_sum(self, t):
        got = self._exchange([t.clone() for _ in range(self.N)])
        t.copy_(torch.stack(got).sum(0))

    def share_ptrs(self, tables):
        # one process, one GPU: the "peer mappings" are the other threads' tensors themselves
        got = self._exchange([[int(t.data_ptr()) for t in tables] for _ in range(self.N)])
        return got
