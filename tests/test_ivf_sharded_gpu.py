"""IVF list shards behind the C ABI (cm_ivf_sharded_*): one process, the inverted lists spread over several shards
(on the one-GPU test box they share the device; with more GPUs visible they spread over them).  Every answer is
compared bit-exactly with the oracle's SINGLE IVFIndex on the same centroids: the cross-shard merge must reproduce the
reference's tie order -- the candidate's number in the append loop over all probed lists (ivf_index_search.go:252-308)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from comet_b200 import capi  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from tests.parity import assert_same_results  # noqa: E402


def _devices(n_shards):
    import torch
    n_dev = torch.cuda.device_count()
    return [r % n_dev for r in range(n_shards)]


def build(n, d, nlist, metric, seed, shards, dup=True):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    if metric == capi.COSINE:
        x += 0.3
    if dup:                                      # identical rows: the same score in different lists -> ties across shards
        x[n // 2] = x[3]
        x[n - 7] = x[3]
        x[100:140] = x[40:80]
    ids = np.arange(1, n + 1, dtype=np.uint32)
    o = O.IVF(d, nlist, metric)
    o.train(x[:max(nlist * 4, 500)].copy())
    g = capi.ShardedIVFIndex(d, nlist, metric, _devices(shards))
    g.set_centroids(o.centroids())
    o.add(ids, x.copy())
    lists = g.add(ids, x.copy())
    return g, o, rng, x, lists


def check(g, o, q, k, nprobes, **kw):
    ids, sc, cnt = g.search(q, k=k, nprobes=nprobes, **kw)
    for i in range(q.shape[0]):
        oi, os_ = o.search(q[i], k=k, nprobes=nprobes, **kw)
        assert_same_results(ids[i], sc[i], cnt[i], oi, os_, what=f"query {i}")


@pytest.mark.parametrize("metric", [capi.L2, capi.L2SQ, capi.COSINE])
def test_list_shards_match_the_single_index(metric):
    g, o, rng, x, lists = build(6000, 32, 24, metric, 100 + metric, 4)
    ol = o.lists()
    for l in range(24):                          # same membership as the single index, list by list
        assert np.array_equal((np.nonzero(lists == l)[0] + 1).astype(np.uint32), ol[l][0])
    assert sum(g.shard_size(r) for r in range(4)) == 6000 and len(g) == 6000
    assert [g.owner(l) for l in range(6)] == [0, 1, 2, 3, 0, 1]
    q = rng.standard_normal((9, 32)).astype(np.float32)
    q[0], q[1] = x[3], x[45]
    for nprobes in (1, 5, 24):
        check(g, o, q, 10, nprobes)
    check(g, o, q, 150, 8)
    check(g, o, q[:3], 0, 2)                     # k <= 0: every candidate of the probed lists
    check(g, o, q[:3], 40, 6, threshold=30.0 if metric != capi.COSINE else 0.9)
    check(g, o, q[:3], 40, 6, filter_ids=np.arange(2, 6000, 3, dtype=np.uint32))


def test_list_shards_delete_flush_rebalance_and_more_adds():
    g, o, rng, x, lists = build(5000, 24, 16, capi.L2SQ, 7, 3)
    q = rng.standard_normal((6, 24)).astype(np.float32)
    q[0] = x[3]
    for dead in (4, 2501, 4994, 77):
        g.remove(dead)
        o.remove(dead)
    with pytest.raises(capi.CometError) as e:
        g.remove(77)
    assert e.value.code == capi.ERR_NOT_FOUND and "already deleted" in e.value.msg
    with pytest.raises(capi.CometError) as e:
        g.remove(999999)
    assert e.value.code == capi.ERR_NOT_FOUND
    check(g, o, q, 20, 5)
    g.flush()
    o.flush()
    assert len(g) == 4996
    check(g, o, q, 20, 5)
    before = [g.shard_size(r) for r in range(3)]
    g.rebalance()                                # greedy by length: lists move whole, answers do not change
    after = [g.shard_size(r) for r in range(3)]
    assert sum(after) == 4996 and max(after) - min(after) <= max(before) - min(before)
    check(g, o, q, 20, 5)
    check(g, o, q, 20, 16)
    more = np.arange(9001, 9301, dtype=np.uint32)
    xm = rng.standard_normal((300, 24)).astype(np.float32)
    g.add(more, xm.copy())
    o.add(more, xm.copy())
    check(g, o, q, 30, 7)
    # per-shard scanned vectors of the last search add up to the single index's scan
    g.search(q, k=10, nprobes=7)
    scanned = g.last_scanned()
    total = 0
    for i in range(len(q)):
        o.search(q[i], k=10, nprobes=7)
    assert scanned.sum() > 0 and len(scanned) == 3


def test_list_shards_errors():
    d = 16
    g = capi.ShardedIVFIndex(d, 4, capi.COSINE, _devices(2))
    x = np.random.default_rng(1).standard_normal((200, d)).astype(np.float32) + 0.2
    with pytest.raises(capi.CometError) as e:
        g.add(np.arange(1, 11, dtype=np.uint32), x[:10].copy())
    assert e.value.code == capi.ERR_NOT_TRAINED
    with pytest.raises(capi.CometError) as e:
        g.search(x[:2], k=3)
    assert e.value.code == capi.ERR_NOT_TRAINED
    g.train(x.copy())
    o = O.IVF(d, 4, capi.COSINE)
    o.set_centroids(g.centroids())
    ids, sc, cnt = g.search(x[:2], k=3)
    assert cnt.tolist() == [0, 0]
    xb = x[:20].copy()
    xb[12] = 0
    with pytest.raises(capi.CometError) as e:     # a zero row stops the batch at that row, like n successive Adds
        g.add(np.arange(1, 21, dtype=np.uint32), xb)
    assert e.value.code == capi.ERR_ZERO_VECTOR and len(g) == 12
    o.add(np.arange(1, 13, dtype=np.uint32), x[:12].copy())
    check(g, o, x[30:33].copy(), 5, 4)
    qz = x[:2].copy()
    qz[1] = 0
    with pytest.raises(capi.CometError) as e:
        g.search(qz, k=3)
    assert e.value.code == capi.ERR_ZERO_VECTOR
    with pytest.raises(capi.CometError) as e:
        g.search(np.ones((1, d + 1), np.float32), k=3)
    assert e.value.code == capi.ERR_DIM_MISMATCH


# ---- IVFPQ list shards ---------------------------------------------------------------------------------------------
def build_pq(n, d, nlist, M, nbits, metric, seed, shards):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32) + (0.2 if metric == capi.COSINE else 0.0)
    x[n // 2] = x[3]                              # identical rows -> identical codes in the same list; ADC scores tie anyway
    ids = np.arange(1, n + 1, dtype=np.uint32)
    o = O.IVFPQ(d, metric, nlist, M, nbits)
    o.train(x[:max(nlist * 10, (1 << nbits) * 2, 800)].copy())
    g = capi.ShardedIVFPQIndex(d, metric, nlist, M, nbits, _devices(shards))
    g.set_trained(o.centroids(), o.codebooks())
    o.add(ids, x.copy())
    lists = g.add(ids, x.copy())
    return g, o, rng, x, lists


@pytest.mark.parametrize("metric", [capi.L2, capi.COSINE])
def test_ivfpq_list_shards_match_the_single_index(metric):
    g, o, rng, x, lists = build_pq(7000, 32, 18, 8, 4, metric, 200 + metric, 4)     # 16 codewords: heavy score ties
    ol = o.lists()
    for l in range(18):
        assert np.array_equal((np.nonzero(lists == l)[0] + 1).astype(np.uint32), ol[l][0])
    assert len(g) == 7000 and sum(g.shard_size(r) for r in range(4)) == 7000
    q = rng.standard_normal((8, 32)).astype(np.float32)
    q[0] = x[3]
    for nprobes in (1, 4, 18):
        check(g, o, q, 10, nprobes)
    check(g, o, q, 200, 6)
    check(g, o, q[:3], 0, 2)
    check(g, o, q[:3], 50, 5, filter_ids=np.arange(2, 7000, 4, dtype=np.uint32))
    for dead in (5, 3500, 6999):
        g.remove(dead)
        o.remove(dead)
    check(g, o, q, 30, 6)
    g.flush()
    o.flush()
    check(g, o, q, 30, 6)
    g.rebalance()
    assert sum(g.shard_size(r) for r in range(4)) == 6997
    check(g, o, q, 30, 6)
    check(g, o, q, 30, 18)
    more = np.arange(9001, 9201, dtype=np.uint32)
    xm = rng.standard_normal((200, 32)).astype(np.float32) + (0.2 if metric == capi.COSINE else 0.0)
    g.add(more, xm.copy())
    o.add(more, xm.copy())
    check(g, o, q, 30, 7)
    g.search(q, k=10, nprobes=7)
    assert g.last_scanned().sum() > 0


def test_ivfpq_list_shards_train_on_device_and_errors():
    rng = np.random.default_rng(5)
    d = 16
    x = rng.standard_normal((1500, d)).astype(np.float32)
    g = capi.ShardedIVFPQIndex(d, capi.L2, 6, 4, 4, _devices(3))
    with pytest.raises(capi.CometError) as e:
        g.search(x[:2], k=3)
    assert e.value.code == capi.ERR_NOT_TRAINED
    g.train(x.copy())
    o = O.IVFPQ(d, capi.L2, 6, 4, 4)
    o.set_trained(*g.trained_state())
    ids = np.arange(1, 1501, dtype=np.uint32)
    g.add(ids, x.copy())
    o.add(ids, x.copy())
    check(g, o, x[7:12].copy(), 15, 3)
    with pytest.raises(capi.CometError):
        capi.ShardedIVFPQIndex(d, capi.L2, 6, 5, 4, _devices(2))       # dim not divisible by M
