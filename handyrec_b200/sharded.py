"""Row-sharded embedding tables: owner(r) = r % N, local row r // N  (SURVEY.md §8e).

The reference is single-device; this is the multi-GPU face of the same lookup.  One process per GPU.
Per step and per rank (requester = the rank that owns the SAMPLE, owner = the rank that owns the ROW):

  forward   route ids (owner + owner-side key per position, grouped by owner)      [hrb_route_ids]
            all-to-all #0  counts            (N int64, equal split)     -> ONE host sync for the split sizes
            all-to-all #1  keys              (4 B per valid position)
            owner gathers its rows                                              [hrb_rows_by_key]
            all-to-all #2  rows              (4*D B per valid position)
            requester scatters / pools them into the group's output block       [hrb_scatter_rows]
  backward  requester gathers per-position gradients in the same order          [hrb_gather_grads]
            all-to-all #3  gradient rows
            owner: sort -> segment-reduce -> row update of its shard            [hrb_keyed_bwd_update]
  dense parameters are replicated; their flat gradient buffer is all-reduced once.

`RowExchange` holds the protocol; the arithmetic lives behind a small provider object so that the protocol can
be exercised under gloo on CPU tensors by the tests (tests/ supplies an oracle-backed provider; the product has
only the CUDA provider below -- there is no CPU fallback in this package).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from . import kernels as K
from ._lib import call
from .engine import DeepFMEngine


def shard_rows(vocab: int, rank: int, n_ranks: int) -> int:
    """Rows of a `vocab`-row table held by `rank` (at least 1 so that every shard is a valid table)."""
    return max(1, (vocab - rank + n_ranks - 1) // n_ranks)


def key_base_table(vocabs: Sequence[int], n_ranks: int) -> torch.Tensor:
    """key_base[r, t] = first key of table t in rank r's shard key space (uint32 stored as int64 on host)."""
    kb = torch.zeros(n_ranks, len(vocabs), dtype=torch.int64)
    for r in range(n_ranks):
        acc = 0
        for t, v in enumerate(vocabs):
            kb[r, t] = acc
            acc += shard_rows(v, r, n_ranks)
    return kb


class CudaShardProvider:
    """The arithmetic of the exchange on this rank's GPU (C ABI calls only)."""

    def __init__(self, plan: K.LookupPlan, vocabs: Sequence[int], n_ranks: int, batch: int):
        self.plan, self.N, self.dev = plan, n_ranks, plan.device
        self.D = plan.tables[0].shape[1]
        self.kb = key_base_table(vocabs, n_ranks).to(torch.int32).to(self.dev)  # values < 2^31
        need = ctypes.c_size_t(0)
        call("hrb_route_workspace", plan._h, batch, ctypes.byref(need))
        self.ws_route = torch.empty(need.value, device=self.dev, dtype=torch.uint8)
        n = batch * plan.pos_cols
        self.perm = torch.empty(max(n, 1), device=self.dev, dtype=torch.int32)
        self.send_keys = torch.empty(max(n, 1), device=self.dev, dtype=torch.int32)
        self.counts = torch.zeros(n_ranks + 1, device=self.dev, dtype=torch.int64)
        self.pos_rows = None if all(f[1] == 1 and f[2] in ("none", None) for f in plan.fields) else torch.zeros(max(n, 1), self.D, device=self.dev)
        self._ws_bwd = None

    def route(self, ids: torch.Tensor):
        B = ids.shape[0]
        call("hrb_route_ids", self.plan._h, K._p(ids), ids.stride(0), B, self.N, K._p(self.kb), K._p(self.perm), K._p(self.send_keys),
             K._p(self.counts), K._p(self.ws_route), self.ws_route.numel(), K._stream())
        return self.perm, self.send_keys, self.counts

    def rows_by_key(self, keys: torch.Tensor, n: int) -> torch.Tensor:
        out = torch.empty(max(n, 1), self.D, device=self.dev, dtype=torch.float32)
        call("hrb_rows_by_key", self.plan._h, K._p(keys), n, K._p(out), K._stream())
        return out[:n]

    def scatter_rows(self, ids, perm, n, rows, out):
        call("hrb_scatter_rows", self.plan._h, K._p(ids), ids.stride(0), ids.shape[0], K._p(perm), n, K._p(rows), K._p(out), out.stride(0),
             K._p(self.pos_rows), K._stream())

    def gather_grads(self, ids, perm, n, dout) -> torch.Tensor:
        send = torch.empty(max(n, 1), self.D, device=self.dev, dtype=torch.float32)
        call("hrb_gather_grads", self.plan._h, K._p(ids), ids.stride(0), K._p(perm), n, K._p(dout), dout.stride(0), K._p(send), K._stream())
        return send[:n]

    def keyed_update(self, keys, grads, n, op: _lib.OptParams):
        need = ctypes.c_size_t(0)
        call("hrb_keyed_bwd_workspace", self.plan._h, n, ctypes.byref(need))
        if self._ws_bwd is None or self._ws_bwd.numel() < need.value:
            self._ws_bwd = torch.empty(int(need.value * 1.25) + 1024, device=self.dev, dtype=torch.uint8)
        call("hrb_keyed_bwd_update", self.plan._h, K._p(keys), K._p(grads), n, ctypes.byref(op), K._p(self._ws_bwd), self._ws_bwd.numel(), K._stream())


class TorchDistComm:
    """torch.distributed plumbing (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist

        self.dist, self.group = dist, group
        self.N, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def all_to_all_equal(self, send: torch.Tensor) -> torch.Tensor:
        recv = torch.empty_like(send)
        self.dist.all_to_all_single(recv, send, group=self.group)
        return recv

    def all_to_all_v(self, send: torch.Tensor, send_counts: List[int], recv_counts: List[int]) -> torch.Tensor:
        recv = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), device=send.device, dtype=send.dtype)
        self.dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=send_counts, group=self.group)
        return recv

    def all_reduce_sum(self, t: torch.Tensor) -> None:
        self.dist.all_reduce(t, group=self.group)

    def alloc_tables(self, vocabs: Sequence[int], dim: int, device) -> Tuple[List[torch.Tensor], Optional[List[List[int]]]]:
        """Allocate this rank's shards in ONE symmetric-memory arena (torch.distributed._symmetric_memory: CUDA VMM
        allocations exported to every rank of the node and mapped over NVLink) -> (tables, peer_ptrs[r][t]).
        With the gloo backend (CPU tests) plain tensors are returned and peer_ptrs is None."""
        rows = [shard_rows(v, self.rank, self.N) for v in vocabs]
        if self.dist.get_backend(self.group) != "nccl":
            return [torch.zeros(r, dim) for r in rows], None
        import torch.distributed._symmetric_memory as symm_mem

        max_rows = [shard_rows(v, 0, self.N) for v in vocabs]  # identical layout on every rank
        offs, off = [], 0
        for r in max_rows:
            offs.append(off)
            off += (r * dim + 63) // 64 * 64  # 256-byte aligned table starts
        arena = symm_mem.empty((off,), dtype=torch.float32, device=device)
        hdl = symm_mem.rendezvous(arena, group=self.group if self.group is not None else self.dist.group.WORLD)
        self._arena, self._arena_hdl = arena, hdl  # keep the mapping alive
        tables = [arena[o : o + r * dim].view(r, dim) for o, r in zip(offs, rows)]
        peer_ptrs = [[int(hdl.buffer_ptrs[rk]) + o * 4 for o in offs] for rk in range(self.N)]
        return tables, peer_ptrs


class RowExchange:
    """The requester/owner protocol of one rank.  `provider` does the arithmetic, `comm` moves the bytes."""

    def __init__(self, provider, comm):
        self.p, self.comm = provider, comm
        self.N = comm.N

    def route_async(self, ids: torch.Tensor) -> None:
        """Launch the routing of `ids` (owner + key per position, grouped by owner) and the exchange of the per-rank counts;
        nothing waits on the host.  Routing depends on the ids only, so it can run long before its results are needed."""
        self.perm, self.send_keys, counts = self.p.route(ids)
        self._counts_dev = counts
        self._recv_counts_dev = self.comm.all_to_all_equal(counts[: self.N].contiguous())
        self._routed = False

    def finish_route(self) -> None:
        if self._routed:
            return
        both = torch.cat([self._counts_dev[: self.N], self._recv_counts_dev]).tolist()  # the one host sync of the step
        self.send_counts, self.recv_counts = [int(x) for x in both[: self.N]], [int(x) for x in both[self.N :]]
        self.n_send, self.n_recv = sum(self.send_counts), sum(self.recv_counts)
        self.recv_keys = self.comm.all_to_all_v(self.send_keys[: self.n_send], self.send_counts, self.recv_counts)
        self._routed = True

    def forward(self, ids: torch.Tensor, out: torch.Tensor) -> None:
        self.route_async(ids)
        self.finish_route()
        rows = self.p.rows_by_key(self.recv_keys, self.n_recv)
        got = self.comm.all_to_all_v(rows, self.recv_counts, self.send_counts)
        self.p.scatter_rows(ids, self.perm, self.n_send, got, out)

    def backward_update(self, ids: torch.Tensor, dout: torch.Tensor, op) -> None:
        self.finish_route()
        send = self.p.gather_grads(ids, self.perm, self.n_send, dout)
        recv = self.comm.all_to_all_v(send, self.send_counts, self.recv_counts)
        self.p.keyed_update(self.recv_keys, recv, self.n_recv, op)


class ShardedDeepFMEngine(DeepFMEngine):
    """DeepFMEngine with every embedding table row-sharded over the ranks of `comm`.

    `tables` are this rank's SHARDS (rows r with r % N == rank, in order); `vocabs` the full vocabulary sizes.
    The per-rank batch is fixed (weak scaling); gradients are averaged over the global batch.
    """

    def __init__(self, tables, vocabs, fields, n_dense, comm, peer_ptrs=None, **kw):
        super().__init__(tables, fields, n_dense, **kw)
        self.comm = comm
        self.world = comm.N
        self.provider = CudaShardProvider(self.plan, vocabs, comm.N, self.B)
        self.exchange = RowExchange(self.provider, comm)
        self.grad_scale_div = comm.N
        self.overlap_embedding_bwd = True
        self.peer_lookup = False
        self._emb_pending = None
        self._barrier_buf = torch.zeros(1, device=self.dev)
        if peer_ptrs is not None and all(f[1] == 1 and f[2] in ("none", None) for f in fields):
            # forward without any all-to-all: every rank maps every other rank's shards (comm.alloc_tables: symmetric memory over
            # NVLink) and the fused lookup+FM kernel gathers row `id` from rank id % N at local row id // N
            n_t = len(self.tables)
            ptrs = (ctypes.c_void_p * (comm.N * n_t))(*[int(peer_ptrs[r][t]) for r in range(comm.N) for t in range(n_t)])
            full = (ctypes.c_int64 * n_t)(*[int(v) for v in vocabs])
            call("hrb_plan_set_peers", self.plan._h, comm.N, ptrs, full)
            self.peer_lookup = True

    def _lookup_fm_forward(self, ids, B, st):
        if self.peer_lookup:
            # the routing needed by the BACKWARD exchange depends on the ids only: launch it now, beside the forward
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream()
            if self._marks is None:
                self._side.wait_stream(main)
                with torch.cuda.stream(self._side):
                    self.exchange.route_async(ids)
            else:
                self.exchange.route_async(ids)
            super()._lookup_fm_forward(ids, B, st)
            return
        self.exchange.forward(ids, self.X0)  # the plan's out_col already includes the dense block
        self._mark("sharded_lookup_fwd")
        emb = self.X0[:, self.nd_pad :]
        call("hrb_fm_fwd", K._p(emb), self.K0p, B, self.F, self.D, K._p(self.fm_w), K._p(self.fm_w0), K._p(self.fm_out), K._p(self.fm_sum), st)
        self._mark("fm_fwd")

    def _pre_embedding_backward(self):
        if self.peer_lookup:
            self.exchange.finish_route()

    def _embedding_backward(self, ids, B, st, op):
        self.exchange.backward_update(ids, self.dX0, op)
        if self.peer_lookup:
            # peers read this shard in the next forward: nobody may start it before every rank has finished updating
            self.comm.all_reduce_sum(self._barrier_buf)
        self._mark("sharded_embedding_bwd_update")

    def _sync_dense_grads(self):
        self.comm.all_reduce_sum(self.grads)
        self._mark("dense_allreduce")
