"""GPU parity tests of the HNSW search path against the CPU oracle on the SAME graph (the oracle builds it
with the reference's insertNode / selectNeighbors / pruneConnections and a seeded level draw; the device
loads it): ids, ranks, score bits AND the work counters (distance evaluations, expansions) must match --
the heaps are Go's container/heap replayed operation by operation."""
import numpy as np
import pytest

from comet_b200 import capi
from oracle import oracle_py as O
from tests.parity import assert_same_results, bits

pytestmark = pytest.mark.gpu


def build_pair(n, d, metric, m, efc, efs, seed, first_level=None):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    levels = O.hnsw_random_levels(n, m, seed)
    if first_level is not None:
        levels[0] = first_level          # the first node is the entry point forever (SURVEY quirk b)
    o = O.HNSW(d, metric, m, efc, efs)
    o.add(ids, x.copy(), levels)
    eids, elev, erows, layers = o.export()
    g = capi.HNSWIndex(d, metric, m, efc, efs)
    g.load_graph(eids, erows, elev, layers, o.entry_point, o.max_level)
    return g, o, rng


def check(g, o, q, k, ef=0, **kw):
    ids, sc, cnt, work = g.search(q, k=k, ef_search=ef, with_work=True, **kw)
    for i in range(q.shape[0]):
        oi, os_ = o.search(q[i], k=k, ef_search=ef, **kw)
        assert_same_results(ids[i], sc[i], cnt[i], oi, os_, what=f"query {i}")
        assert tuple(work[i]) == O.HNSW.last_counters(), f"query {i}: work counters differ"


@pytest.mark.parametrize("metric", [capi.L2, capi.L2SQ, capi.COSINE])
def test_hnsw_search_matches_oracle(metric):
    g, o, rng = build_pair(3000, 32, metric, 8, 60, 40, 10 + metric, first_level=2)
    assert o.max_level >= 2
    q = rng.standard_normal((19, 32)).astype(np.float32)
    check(g, o, q, 10)
    check(g, o, q, 10, ef=100)
    check(g, o, q[:5], 0, ef=64)             # WithK(0): every surviving candidate
    check(g, o, q[:5], 1000, ef=16)          # k > ef


def test_hnsw_entry_at_level0_dim768_m16():
    # the common case of the reference: the first node drew level 0, so phase 1 is a no-op
    g, o, rng = build_pair(1500, 768, capi.COSINE, 16, 80, 128, 3, first_level=0)
    q = rng.standard_normal((8, 768)).astype(np.float32)
    check(g, o, q, 10)


def test_hnsw_filter_threshold_delete():
    g, o, rng = build_pair(2500, 24, capi.L2, 8, 60, 50, 21, first_level=1)
    q = rng.standard_normal((9, 24)).astype(np.float32)
    check(g, o, q, 10, threshold=5.0)
    check(g, o, q, 10, filter_ids=np.arange(1, 2501, 2, dtype=np.uint32))
    for i in range(5, 900, 6):
        g.remove(i)
        o.remove(i)
    check(g, o, q, 10)
    g.remove(1)                               # the entry point: searchLayer then returns nothing
    o.remove(1)
    check(g, o, q, 10)


def test_hnsw_empty_and_errors():
    g = capi.HNSWIndex(8, capi.L2)
    assert g.search(np.ones((2, 8), np.float32), k=3)[2].tolist() == [0, 0]
    with pytest.raises(capi.CometError) as e:
        g.search(np.ones((1, 9), np.float32), k=1)
    assert e.value.code == capi.ERR_DIM_MISMATCH
    assert capi.HNSWIndex(8, capi.L2, 0, 0, 0).ef_search() == 200   # hnsw_index.go:178-190 defaults


def oracle_flat_graph(o):
    """(slot, layer)-ordered neighbour-ID lists of the oracle's graph, as cm_hnsw_export_graph lays them out."""
    eids, elev, erows, layers = o.export()
    off = [0]
    chunks = []
    for s in range(len(eids)):
        for layer in range(int(elev[s]) + 1):
            lo, nb = layers[layer]
            seg = nb[lo[s]:lo[s + 1]]
            chunks.append(seg)
            off.append(off[-1] + len(seg))
    flat = np.concatenate(chunks) if chunks else np.zeros(0, np.uint32)
    return elev, np.asarray(off, np.int64), flat.astype(np.uint32)


@pytest.mark.parametrize("metric", [capi.L2, capi.COSINE])
def test_hnsw_device_insertion_builds_the_oracles_graph(metric):
    # HNSWIndex.Add / insertNode / selectNeighbors / pruneConnections on the device, same level draws
    rng = np.random.default_rng(31 + metric)
    n, d, m, efc = 1200, 24, 6, 40
    x = rng.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    levels = O.hnsw_random_levels(n, m, 5)
    levels[0] = 1
    o = O.HNSW(d, metric, m, efc, 50)
    o.add(ids, x.copy(), levels)
    g = capi.HNSWIndex(d, metric, m, efc, 50)
    xg = x.copy()
    for a, b in [(0, 1), (1, 400), (400, 401), (401, n)]:          # several Add batches, incl. single nodes
        g.add(ids[a:b], xg[a:b], levels[a:b])
    assert len(g) == n and g.max_level() == o.max_level
    if metric == capi.COSINE:                                       # Add normalised the caller's rows in place
        assert np.array_equal(bits(xg), bits(np.stack([O.normalize(r) for r in x])))
    glev, goff, gids, gentry, gml = g.export_graph()
    olev, ooff, oflat = oracle_flat_graph(o)
    assert gentry == o.entry_point and gml == o.max_level
    assert np.array_equal(glev, olev)
    assert np.array_equal(goff, ooff), "degrees differ"
    assert np.array_equal(gids, oflat), "edges differ"
    q = rng.standard_normal((11, d)).astype(np.float32)
    check(g, o, q, 10)
    check(g, o, q, 5, ef=30)
    with pytest.raises(capi.CometError):
        g.add([5], x[:1].copy(), [0])                               # duplicate ID
    with pytest.raises(capi.CometError):
        g.add([0], x[:1].copy(), [0])                               # ID 0 is reserved
