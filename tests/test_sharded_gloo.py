"""N > 1 host logic on CPU: two gloo ranks, each owning a contiguous row shard.  The local search and the
merge are stand-ins built on the CPU oracle (the CUDA calls need a GPU); what is tested is the product's
sharding / gather plumbing (comet_b200.sharded) and the merge RULE the device kernel implements:
global order (score, scan position) == (score, shard, rank within the shard's list)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from comet_b200.sharded import ShardedSearch, shard_bounds

N, D, K, NQ = 3001, 24, 10, 7


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _data():
    rng = np.random.default_rng(123)
    x = rng.standard_normal((N, D)).astype(np.float32)
    x[100] = x[2900]            # an exact tie that straddles the two shards
    x[5] = x[6]                 # and one inside a shard
    q = rng.standard_normal((NQ, D)).astype(np.float32)
    q[0] = x[100]
    return x, q


def _merge_rule(g_ids, g_sc, g_cnt):
    """numpy statement of cm_merge_shards_device's rule."""
    world, nq, stride = g_ids.shape
    out_ids = torch.zeros((nq, K), dtype=g_ids.dtype)
    out_sc = torch.zeros((nq, K), dtype=g_sc.dtype)
    out_cnt = torch.zeros((nq,), dtype=torch.int64)
    for qi in range(nq):
        cand = []
        for r in range(world):
            for j in range(int(g_cnt[r, qi])):
                cand.append((float(g_sc[r, qi, j]), r, j, int(g_ids[r, qi, j])))
        cand.sort(key=lambda t: (t[0], t[1], t[2]))
        cand = cand[:K]
        out_cnt[qi] = len(cand)
        for i, c in enumerate(cand):
            out_sc[qi, i] = c[0]
            out_ids[qi, i] = c[3]
    return out_ids, out_sc, out_cnt


def _worker(rank, world, port, metric, out_q):
    from oracle import oracle_py as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, q = _data()
    row0, rows = shard_bounds(N, world, rank)
    o = O.Flat(D, metric)
    o.add(np.arange(row0 + 1, row0 + rows + 1, dtype=np.uint32), x[row0:row0 + rows].copy())

    def search_local(queries):
        ids, sc, cnt = o.search_batch(queries, K)
        return torch.from_numpy(ids.astype(np.int64)), torch.from_numpy(sc), torch.from_numpy(cnt)

    s = ShardedSearch(search_local, _merge_rule)
    ids, sc, cnt = s.search(q)
    if rank == 0:
        out_q.put((ids.numpy(), sc.numpy(), cnt.numpy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("metric", [1, 2])   # l2_squared, cosine
def test_two_rank_sharded_search_matches_single_index(metric):
    from oracle import oracle_py as O
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, metric, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0, "a gloo rank failed"
    ids, sc, cnt = out_q.get(timeout=10)
    x, q = _data()
    whole = O.Flat(D, metric)
    whole.add(np.arange(1, N + 1, dtype=np.uint32), x.copy())
    for i in range(NQ):
        oi, os_ = whole.search(q[i], k=K)
        assert cnt[i] == len(oi)
        assert np.array_equal(sc[i, :cnt[i]].view(np.uint32), os_.view(np.uint32))
        assert np.array_equal(ids[i, :cnt[i]], oi.astype(np.int64)), f"query {i}"


def test_shard_bounds_cover_rows_contiguously():
    for n, w in [(10, 3), (1_000_000, 8), (7, 8), (128, 1)]:
        nxt = 0
        for r in range(w):
            r0, m = shard_bounds(n, w, r)
            assert r0 == nxt and m >= 0
            nxt = r0 + m
        assert nxt == n


# ---- the other layouts of SURVEY 8e on CPU: PQ row shards (same exchange step as flat) and query-sharded replicas ----
def _pq_worker(rank, world, port, out_q):
    from oracle import oracle_py as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, q = _data()
    train = x[:600]
    row0, rows = shard_bounds(N, world, rank)
    o = O.PQ(D, 0, 4, 4)
    o.train(train.copy())                                    # codebooks are replicated: same training set on every rank
    o.add(np.arange(row0 + 1, row0 + rows + 1, dtype=np.uint32), x[row0:row0 + rows].copy())

    def search_local(queries):
        ids = np.zeros((len(queries), K), np.int64)
        sc = np.zeros((len(queries), K), np.float32)
        cnt = np.zeros(len(queries), np.int64)
        for i, qq in enumerate(queries):
            a, b = o.search(qq, k=K)
            ids[i, :len(a)], sc[i, :len(a)], cnt[i] = a, b, len(a)
        return torch.from_numpy(ids), torch.from_numpy(sc), torch.from_numpy(cnt)

    ids, sc, cnt = ShardedSearch(search_local, _merge_rule).search(q)
    if rank == 0:
        out_q.put((ids.numpy(), sc.numpy(), cnt.numpy()))
    dist.destroy_process_group()


def test_two_rank_pq_row_shards_match_single_index():
    """ADC scores tie often (few distinct table sums): the (score, shard, rank-in-list) merge must reproduce the
    single index's (score, position) order across the shard boundary."""
    from oracle import oracle_py as O
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pq_worker, args=(r, 2, port, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0, "a gloo rank failed"
    ids, sc, cnt = out_q.get(timeout=10)
    x, q = _data()
    whole = O.PQ(D, 0, 4, 4)
    whole.train(x[:600].copy())
    whole.add(np.arange(1, N + 1, dtype=np.uint32), x.copy())
    for i in range(NQ):
        oi, os_ = whole.search(q[i], k=K)
        assert cnt[i] == len(oi)
        assert np.array_equal(sc[i, :cnt[i]].view(np.uint32), os_.view(np.uint32))
        assert np.array_equal(ids[i, :cnt[i]], oi.astype(np.int64)), f"query {i}"


def _replica_worker(rank, world, port, out_q):
    from oracle import oracle_py as O
    from comet_b200.sharded import ReplicatedSearch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, q = _data()
    o = O.IVF(D, 8, 0)
    o.train(x[:400].copy())
    o.add(np.arange(1, N + 1, dtype=np.uint32), x.copy())

    def search_local(queries):
        ids = np.zeros((len(queries), K), np.int64)
        sc = np.zeros((len(queries), K), np.float32)
        cnt = np.zeros(len(queries), np.int64)
        for i, qq in enumerate(queries):
            a, b = o.search(qq, k=K, nprobes=3)
            ids[i, :len(a)], sc[i, :len(a)], cnt[i] = a, b, len(a)
        return torch.from_numpy(ids), torch.from_numpy(sc), torch.from_numpy(cnt)

    ids, sc, cnt = ReplicatedSearch(search_local).search(torch.from_numpy(q).numpy())
    if rank == 1:                                            # every rank holds the full result
        out_q.put((ids.numpy(), sc.numpy(), cnt.numpy()))
    dist.destroy_process_group()


def test_three_rank_replicas_answer_their_own_queries():
    from oracle import oracle_py as O
    from comet_b200.sharded import query_bounds
    covered = []
    for r in range(3):
        q0, m = query_bounds(NQ, 3, r)
        covered += list(range(q0, q0 + m))
    assert covered == list(range(NQ))
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_replica_worker, args=(r, 3, port, out_q)) for r in range(3)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0, "a gloo rank failed"
    ids, sc, cnt = out_q.get(timeout=10)
    x, q = _data()
    o = O.IVF(D, 8, 0)
    o.train(x[:400].copy())
    o.add(np.arange(1, N + 1, dtype=np.uint32), x.copy())
    assert ids.shape == (NQ, K)
    for i in range(NQ):
        oi, os_ = o.search(q[i], k=K, nprobes=3)
        assert cnt[i] == len(oi)
        assert np.array_equal(sc[i, :cnt[i]].view(np.uint32), os_.view(np.uint32))
        assert np.array_equal(ids[i, :cnt[i]], oi.astype(np.int64)), f"query {i}"
