"""Row-sharded flat search over the GPUs of one box: one process per GPU (torch.distributed), each
rank owns a contiguous row range of the corpus, searches it with the CUDA path, and the per-shard
top-K lists meet in one exchange step -- an all-gather of [nq][K] (id, score, count) followed by the
device merge `cm_merge_shards_device`.  The reference has no distributed path; what is
preserved is its result: the global order (score, scan position) equals (score, shard, rank within the
shard's list) because shards are contiguous row ranges in scan order.

The collective plumbing is separated from the two device calls so that the N > 1 logic can be
exercised on CPU (gloo) with stand-in callables -- see tests/test_sharded_gloo.py.
"""
from __future__ import annotations

from typing import Callable, Tuple


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Rows [row0, row0 + rows) owned by `rank`: contiguous, in scan order, remainder to the last rank."""
    per = n_rows // world
    row0 = rank * per
    rows = per if rank < world - 1 else n_rows - row0
    return row0, rows


class ShardedSearch:
    """search_local(queries) -> (ids[nq,K], scores[nq,K], counts[nq]) tensors on this rank's device;
    merge(g_ids[W,nq,K], g_scores[W,nq,K], g_counts[W,nq]) -> (ids, scores, counts) global result."""

    def __init__(self, search_local: Callable, merge: Callable, group=None):
        self.search_local = search_local
        self.merge = merge
        self.group = group

    def search(self, queries):
        import torch
        import torch.distributed as dist
        ids, scores, counts = self.search_local(queries)
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        if world == 1:
            return ids, scores, counts
        g_ids = torch.empty((world,) + tuple(ids.shape), dtype=ids.dtype, device=ids.device)
        g_sc = torch.empty((world,) + tuple(scores.shape), dtype=scores.dtype, device=scores.device)
        g_cnt = torch.empty((world,) + tuple(counts.shape), dtype=counts.dtype, device=counts.device)
        # the gathered buffer is the rank-major concatenation along dim 0 ([W*nq, K] viewed as [W, nq, K])
        dist.all_gather_into_tensor(g_ids.view((-1,) + tuple(ids.shape[1:])), ids.contiguous(), group=self.group)
        dist.all_gather_into_tensor(g_sc.view((-1,) + tuple(scores.shape[1:])), scores.contiguous(), group=self.group)
        dist.all_gather_into_tensor(g_cnt.view((-1,) + tuple(counts.shape[1:])), counts.contiguous(), group=self.group)
        return self.merge(g_ids, g_sc, g_cnt)


def device_merge(k: int, stream_ptr: int = 0):
    """The product merge: cm_merge_shards_device on the gathered device tensors."""
    from . import capi

    def merge(g_ids, g_sc, g_cnt):
        import torch
        world, nq, stride = g_ids.shape
        out_ids = torch.empty((nq, k), dtype=g_ids.dtype, device=g_ids.device)
        out_sc = torch.empty((nq, k), dtype=g_sc.dtype, device=g_sc.device)
        out_cnt = torch.empty((nq,), dtype=torch.int64, device=g_ids.device)
        capi.merge_shards_device(g_ids.data_ptr(), g_sc.data_ptr(), g_cnt.data_ptr(), world, nq, stride, k,
                                 out_ids.data_ptr(), out_sc.data_ptr(), out_cnt.data_ptr(), out_stride=k,
                                 stream=stream_ptr)
        return out_ids, out_sc, out_cnt

    return merge


def query_bounds(nq: int, world: int, rank: int) -> Tuple[int, int]:
    """Queries [q0, q0 + m) answered by `rank` when every rank holds a REPLICA of the index: equal contiguous
    shares, remainder spread over the first ranks."""
    per, rem = divmod(nq, world)
    q0 = rank * per + min(rank, rem)
    return q0, per + (1 if rank < rem else 0)


class ReplicatedSearch:
    """Query-sharded search over replicas: the layout for every index whose device state fits one GPU
    (IVF / PQ / IVFPQ codes and lists, the HNSW graph -- SURVEY 8e "replicas only" -- and the flat corpus of
    the headline config).  Queries are the independent units of the path, so there is NO data-path
    collective: rank r answers its own contiguous share with `search_local`, and the result is exactly what
    one index would return, ties included (nothing is merged).  `gather=True` additionally assembles the
    full [nq, K] result on every rank (one all-gather of padded shares) for callers that want it in one place.

    search_local(queries[m, d]) -> (ids[m, K], scores[m, K], counts[m]) tensors."""

    def __init__(self, search_local: Callable, group=None):
        self.search_local = search_local
        self.group = group

    def search(self, queries, gather: bool = True):
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        nq = int(queries.shape[0])
        q0, m = query_bounds(nq, world, rank)
        ids, scores, counts = self.search_local(queries[q0:q0 + m])
        if world == 1 or not gather:
            return ids, scores, counts
        share = (nq + world - 1) // world            # padded share: all_gather needs equal shapes

        def pad(t):
            out = torch.zeros((share,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            out[:m] = t
            return out
        parts = []
        for t in (ids, scores, counts):
            buf = torch.empty((world * share,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(buf, pad(t).contiguous(), group=self.group)
            rows = []
            for r in range(world):
                _, mr = query_bounds(nq, world, r)
                rows.append(buf[r * share:r * share + mr])
            parts.append(torch.cat(rows, dim=0))
        return tuple(parts)
