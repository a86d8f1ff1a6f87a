"""GPU parity tests of the IVF path: CUDA (through the C ABI) vs the CPU oracle on a SHARED trained index
(the oracle trains with the reference's deterministic k-means, the device loads those centroids):
list assignment, ids, ranks and score bits must be identical."""
import numpy as np
import pytest

from comet_b200 import capi
from oracle import oracle_py as O
from tests.parity import assert_same_results

pytestmark = pytest.mark.gpu


def build_pair(n, d, nlist, metric, seed, n_train=None):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    if metric == capi.COSINE:
        x += 0.3          # not centred: cosine lists differ in size
    ids = np.arange(1, n + 1, dtype=np.uint32)
    o = O.IVF(d, nlist, metric)
    o.train(x[: (n_train or n)].copy())
    g = capi.IVFIndex(d, nlist, metric)
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
def test_ivf_assignment_and_search_match_oracle(metric):
    g, o, rng, x, lists = build_pair(6000, 48, 37, metric, 100 + metric)
    # list membership: same lists, same insertion order
    ol = o.lists()
    for l in range(37):
        mine = np.nonzero(lists == l)[0] + 1
        assert np.array_equal(mine.astype(np.uint32), ol[l][0]), f"list {l}"
    q = rng.standard_normal((21, 48)).astype(np.float32)
    for nprobes in (1, 6, 37):
        check(g, o, q, 10, nprobes)
    check(g, o, q[:5], 0, 3)          # WithK(0): every candidate of the probed lists
    check(g, o, q[:5], 100, 0)        # nprobes <= 0 -> all lists


def test_ivf_dim768_default_nprobes():
    g, o, rng, x, _ = build_pair(5000, 768, 64, capi.COSINE, 7)
    assert g.default_nprobes() == o.default_nprobes() == 8
    q = rng.standard_normal((16, 768)).astype(np.float32)
    check(g, o, q, 100, 8)


def test_ivf_threshold_filter_delete_flush():
    g, o, rng, x, _ = build_pair(4000, 32, 20, capi.L2, 9)
    q = rng.standard_normal((9, 32)).astype(np.float32)
    check(g, o, q, 15, 5, threshold=6.5)
    allow = np.arange(1, 4001, 3, dtype=np.uint32)
    check(g, o, q, 15, 5, filter_ids=allow)
    for i in range(2, 900, 5):
        g.remove(i)
        o.remove(i)
    check(g, o, q, 15, 5)
    g.flush()
    o.flush()
    assert len(g) == o.total()
    check(g, o, q, 15, 20)
    extra = rng.standard_normal((50, 32)).astype(np.float32)
    eid = np.arange(5001, 5051, dtype=np.uint32)
    g.add(eid, extra.copy())
    o.add(eid, extra.copy())
    check(g, o, q, 15, 5)


def test_ivf_errors_and_small_reference_cases():
    g = capi.IVFIndex(4, 2, capi.L2)
    with pytest.raises(capi.CometError) as e:
        g.search(np.zeros((1, 4), np.float32), k=1, nprobes=1)
    assert e.value.code == capi.ERR_NOT_TRAINED and "index must be trained before searching" in e.value.msg
    with pytest.raises(capi.CometError) as e:
        g.add([1], np.ones((1, 4), np.float32))
    assert e.value.code == capi.ERR_NOT_TRAINED
    g.set_centroids(np.array([[0, 0, 0, 0], [10, 10, 10, 10]], np.float32))
    assert g.search(np.zeros((1, 4), np.float32), k=3, nprobes=1)[2][0] == 0     # trained but empty
    g.add([1, 2, 3], np.array([[0, 0, 0, 1], [9, 9, 9, 9], [1, 0, 0, 0]], np.float32))
    ids, sc, cnt = g.search(np.zeros((1, 4), np.float32), k=5, nprobes=1)
    assert cnt[0] == 2 and ids[0, :2].tolist() == [1, 3] and sc[0, :2].tolist() == [1.0, 1.0]
    with pytest.raises(capi.CometError) as e:
        g.search(np.zeros((1, 5), np.float32), k=1, nprobes=1)
    assert e.value.code == capi.ERR_DIM_MISMATCH


def test_ivf_k_all_beyond_the_shared_memory_merge():
    # WithK(0) = every candidate of the probed lists (limiter.go:12-17): 30,000 candidates per query here, far more
    # than the shared-memory merge holds -- the sort fallback must give the same order, ties included
    g, o, rng, x, lists = build_pair(30000, 16, 4, capi.L2SQ, 91, n_train=3000)
    q = rng.standard_normal((3, 16)).astype(np.float32)
    q[0] = x[5]
    check(g, o, q, 0, 4)
    check(g, o, q, 20000, 3)
    check(g, o, q[:1], 0, 2, threshold=20.0)
    # and a batch: the list-major scan's parts through the same sort
    qb = rng.standard_normal((34, 16)).astype(np.float32)
    ids, sc, cnt = g.search(qb, k=0, nprobes=3)
    for i in (0, 17, 33):
        oi, os_ = o.search(qb[i], k=0, nprobes=3)
        assert_same_results(ids[i], sc[i], cnt[i], oi, os_, what=f"batch query {i}")


# Batches take the list-major scan (the (query, probe) pairs grouped by list, a 128-row tile of a list walked for up to 8
# queries at once); candidate numbers -- the tie-break key -- and parts per query are laid out differently from the
# query-major scan, the answers must not be.
@pytest.mark.parametrize("metric,d", [(capi.L2, 48), (capi.COSINE, 100), (capi.L2SQ, 768)])
def test_ivf_list_major_scan(metric, d, monkeypatch):
    n = 7000 if d < 768 else 3000
    g, o, rng, x, lists = build_pair(n, d, 23, metric, 500 + d)
    x[100:140] = x[7]                           # ties inside and across lists are ordered by candidate number
    q = rng.standard_normal((70, d)).astype(np.float32) + (0.3 if metric == capi.COSINE else 0.0)
    q[3] = x[7] if metric != capi.COSINE else q[3]
    for nprobes in (1, 5, 23):
        check(g, o, q, 10, nprobes)
    check(g, o, q[:40], 0, 2)                   # WithK(0)
    check(g, o, q[:33], 200, 4, threshold=float(np.sqrt(2 * d)) if metric == capi.L2 else (2.0 * d if metric == capi.L2SQ else 0.9))
    check(g, o, q[:33], 50, 6, filter_ids=np.arange(2, n, 3, dtype=np.uint32))
    for i in range(5, n, 4):
        g.remove(i)
        o.remove(i)
    check(g, o, q[:35], 20, 7)
    g.flush()
    o.flush()
    check(g, o, q[:35], 20, 7)
    # the same batch through both scans, and a small batch forced through the list-major one
    lm = g.search(q, k=30, nprobes=9)
    monkeypatch.setenv("COMET_B200_IVF_LIST_MAJOR", "0")
    qm = g.search(q, k=30, nprobes=9)
    monkeypatch.setenv("COMET_B200_IVF_LIST_MAJOR", "1")
    check(g, o, q[:3], 10, 4)
    check(g, o, q[:1], 0, 23)
    monkeypatch.delenv("COMET_B200_IVF_LIST_MAJOR")
    for a, b in zip(lm, qm):
        assert np.array_equal(a, b)
