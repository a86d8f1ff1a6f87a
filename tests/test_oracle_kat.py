"""Pins the CPU oracle against every known-answer test the reference's own *_test.go files
hold for the hot path.  Each test names the reference test it replays.
CPU-only: runs in the `-m "not gpu"` suite."""
import math

import numpy as np
import pytest

from oracle import oracle_py as O

EPS = 1e-6  # distance_test.go:12 epsilon


def almost(a, b, eps=EPS):
    return abs(a - b) < eps


# ---- distance_test.go:87-145 TestEuclideanCalculate -------------------------------------
@pytest.mark.parametrize("a,b,want", [
    ([1, 2, 3], [1, 2, 3], 0.0),
    ([0, 0], [3, 4], 5.0),
    ([1, 2, 2], [1, 2, 3], 1.0),
    ([-1, -2], [1, 2], 4.472136),
    ([0, 0, 0], [0, 0, 0], 0.0),
    ([5], [2], 3.0),
])
def test_euclidean_calculate(a, b, want):
    assert almost(O.distance(O.L2, a, b), want)


# ---- distance_test.go:214-266 TestL2SquaredCalculate ------------------------------------
@pytest.mark.parametrize("a,b,want", [
    ([1, 2, 3], [1, 2, 3], 0.0),
    ([0, 0], [3, 4], 25.0),
    ([1, 2, 2], [1, 2, 3], 1.0),
    ([-1, -2], [1, 2], 20.0),
    ([0, 0, 0], [0, 0, 0], 0.0),
])
def test_l2squared_calculate(a, b, want):
    assert almost(O.distance(O.L2SQ, a, b), want)


# ---- distance_test.go:335-387 TestCosineCalculate ---------------------------------------
@pytest.mark.parametrize("a,b,want", [
    ([0.6, 0.8], [0.6, 0.8], 0.0),
    ([1, 0], [0, 1], 1.0),
    ([1, 0], [-1, 0], 2.0),
    ([0.707107, 0.707107], [1, 0], 0.292893),
    ([0.5, 0.5, 0.5, 0.5], [0.5, 0.5, 0.5, 0.5], 0.0),
])
def test_cosine_calculate(a, b, want):
    assert almost(O.distance(O.COSINE, a, b), want)


# ---- distance_test.go:417-491 TestCosinePreprocess / InPlace ----------------------------
def test_cosine_preprocess():
    v = np.array([3, 4], np.float32)
    r = O.preprocess(O.COSINE, v)
    assert almost(r[0], 0.6) and almost(r[1], 0.8)
    assert v[0] == 3.0 and v[1] == 4.0           # original untouched by Preprocess
    assert almost(O.norm(r), 1.0)
    with pytest.raises(O.OracleError) as e:
        O.preprocess(O.COSINE, [0, 0, 0])
    assert e.value.code == O.ERR_ZERO_VECTOR     # ErrZeroVector


def test_euclidean_preprocess_is_noop():
    # distance_test.go:147-212: euclidean / l2squared Preprocess returns the vector unchanged
    v = np.array([3, 4, 5], np.float32)
    assert np.array_equal(O.preprocess(O.L2, v), v)
    assert np.array_equal(O.preprocess(O.L2SQ, v), v)


def test_empty_vectors_distance_zero():
    # distance_test.go:927-947
    e = np.empty(0, np.float32)
    assert O.distance(O.L2, e, e) == 0.0
    assert O.distance(O.L2SQ, e, e) == 0.0
    assert O.distance(O.COSINE, e, e) == 1.0 - 0.0


def test_sequential_rounding_is_what_is_restated():
    """The oracle's sum must be the strictly sequential float32 sum (distance.go:158-165), which
    differs in low bits from numpy's pairwise float32 sum and from a float64 sum."""
    rng = np.random.default_rng(7)
    a = rng.standard_normal(768).astype(np.float32)
    b = rng.standard_normal(768).astype(np.float32)
    want = np.float32(0)
    for i in range(768):
        diff = np.float32(a[i] - b[i])
        want = np.float32(want + np.float32(diff * diff))
    got = O.distance(O.L2SQ, a, b)
    assert np.float32(got) == want
    # fused variant (arm64 gc): emulate fma in float64 then round once
    O.set_fma(True)
    try:
        wf = np.float32(0)
        for i in range(768):
            diff = np.float64(np.float32(a[i] - b[i]))
            wf = np.float32(diff * diff + np.float64(wf))   # exact product + one rounding
        assert np.float32(O.distance(O.L2SQ, a, b)) == wf
    finally:
        O.set_fma(False)


# ---- limiter_test.go:7-73 TestSanitizeK --------------------------------------------------
@pytest.mark.parametrize("k,n,want", [(0, 10, 10), (-5, 10, 10), (100, 10, 10), (5, 10, 5),
                                      (10, 10, 10), (5, 0, 0), (0, 0, 0), (1, 10, 1)])
def test_sanitize_k(k, n, want):
    assert O.sanitize_k(k, n) == want


# ---- limiter_test.go:185-256 TestAutocut -------------------------------------------------
@pytest.mark.parametrize("scores,cutoff,want", [
    ([], 1, 0),
    ([1.0], 1, 1),
    ([1.0, 2.0], 1, 2),
    ([0.1, 0.2, 0.3, 0.4, 0.5], 1, 2),
    ([0.1, 0.15, 0.2, 0.5, 0.6, 0.7, 0.8], 1, 3),
    ([0.1, 0.12, 0.13, 0.14, 0.15, 0.8, 0.9, 1.0], 1, 5),
    ([0.1, 0.2, 0.4, 0.45, 0.7, 0.75, 0.9, 1.0], 2, 4),
    ([0.1, 0.2, 0.5, 0.6], 5, 4),
    ([0.5, 0.5, 0.5, 0.5, 0.5], 1, 5),
])
def test_autocut(scores, cutoff, want):
    assert O.autocut(np.array(scores, np.float32), cutoff) == want


# ---- aggregation_test.go:7-157 -----------------------------------------------------------
def test_sum_aggregation_bit_exact():
    ids, sc = O.aggregate("sum", [1, 2, 1, 3, 1], [0.1, 0.2, 0.15, 0.3, 0.05])
    assert len(ids) == 3
    got = dict(zip(ids.tolist(), sc.tolist()))
    assert np.float32(got[1]) == np.float32(0.3)          # exact equality in the reference test
    assert all(sc[i] >= sc[i - 1] for i in range(1, len(sc)))


def test_max_aggregation():
    ids, sc = O.aggregate("max", [1, 2, 1, 1], [0.1, 0.2, 0.5, 0.15])
    assert len(ids) == 2
    assert np.float32(dict(zip(ids.tolist(), sc.tolist()))[1]) == np.float32(0.5)


def test_mean_aggregation_bit_exact():
    ids, sc = O.aggregate("mean", [1, 2, 1, 1], [0.1, 0.2, 0.2, 0.3])
    assert len(ids) == 2
    assert np.float32(dict(zip(ids.tolist(), sc.tolist()))[1]) == np.float32(0.2)


@pytest.mark.parametrize("kind", ["sum", "max", "mean"])
def test_aggregation_empty_and_single(kind):
    ids, sc = O.aggregate(kind, [], [])
    assert len(ids) == 0
    ids, sc = O.aggregate(kind, [1], [0.5])
    assert len(ids) == 1 and sc[0] == np.float32(0.5)


# ---- clustering_test.go:9-57, 203-301 ----------------------------------------------------
def test_kmeans_basic():
    v = np.array([[0, 0], [1, 1], [0.5, 0.5], [10, 10], [11, 11], [10.5, 10.5]], np.float32)
    c, a = O.kmeans(v, 2, O.L2SQ, 20)
    assert c.shape == (2, 2) and len(a) == 6
    assert a[0] == a[1] == a[2] and a[3] == a[4] == a[5] and a[0] != a[3]


def test_kmeans_centroid_accuracy():
    v = np.array([[0, 0], [2, 2], [10, 10], [12, 12]], np.float32)
    c, a = O.kmeans(v, 2, O.L2SQ, 20)
    c0, c1 = a[0], 1 - a[0]
    assert np.allclose(c[c0], [1, 1], atol=0.01) and np.allclose(c[c1], [11, 11], atol=0.01)


def test_kmeans_k_larger_than_n_and_determinism():
    # clustering.go:133-136 auto-adjusts k; clustering.go:147-162 init is deterministic
    v = np.arange(12, dtype=np.float32).reshape(3, 4)
    c, a = O.kmeans(v, 10, O.L2SQ, 20)
    assert c.shape[0] == 3
    rng = np.random.default_rng(3)
    w = rng.standard_normal((200, 8)).astype(np.float32)
    c1, a1 = O.kmeans(w, 7, O.L2SQ, 20)
    c2, a2 = O.kmeans(w, 7, O.L2SQ, 20)
    assert np.array_equal(c1, c2) and np.array_equal(a1, a2)


def test_nearest_centroid_first_min_wins():
    # clustering.go:256-272 strict '<'
    c = np.array([[1, 0], [1, 0], [0, 0]], np.float32)
    assert O.nearest_centroid([1, 0], c, O.L2SQ) == 0


# ---- flat_index_search_test.go -----------------------------------------------------------
def _flat(metric, vecs, first_id=1):
    f = O.Flat(len(vecs[0]), metric)
    rows = np.array(vecs, np.float32)
    f.add(np.arange(first_id, first_id + len(vecs)), rows)
    return f


def test_flat_search_simple():
    # flat_index_search_test.go:10-48
    f = _flat(O.L2, [[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0]])
    ids, sc = f.search([1, 0, 0], k=2)
    assert len(ids) == 2 and ids[0] == 1 and sc[0] == 0.0


def test_flat_search_threshold():
    # flat_index_search_test.go:51-86
    f = _flat(O.L2, [[1, 0, 0], [2, 0, 0], [4, 0, 0], [10, 0, 0]])
    ids, sc = f.search([1, 0, 0], k=10, threshold=2.0)
    assert len(ids) == 2


@pytest.mark.parametrize("k,want", [(0, 5), (-1, 5), (3, 3), (5, 5), (100, 5), (1, 1)])
def test_flat_search_k_bounds(k, want):
    # flat_index_search_test.go:348-389
    f = _flat(O.L2, [[float(i), 0, 0] for i in range(5)])
    ids, _ = f.search([0, 0, 0], k=k)
    assert len(ids) == want


def test_flat_search_results_ordered():
    # flat_index_search_test.go:490-536
    f = _flat(O.L2, [[5, 0, 0], [1, 0, 0], [10, 0, 0], [3, 0, 0]])
    ids, sc = f.search([0, 0, 0], k=4)
    assert sc.tolist() == [1.0, 3.0, 5.0, 10.0]
    assert ids.tolist() == [2, 4, 1, 3]


def test_flat_cosine_add_normalises_in_place_and_zero_vector():
    # flat_index.go:169-189 + F7; flat_index_test.go zero-vector case
    f = O.Flat(2, O.COSINE)
    rows = np.array([[3, 4]], np.float32)
    f.add([1], rows)
    assert almost(rows[0, 0], 0.6) and almost(rows[0, 1], 0.8)   # caller's slice was normalised
    with pytest.raises(O.OracleError):
        f.add([2], np.zeros((1, 2), np.float32))
    ids, sc = f.search([6, 8], k=1)
    assert ids[0] == 1 and almost(sc[0], 0.0)


def test_flat_document_filter():
    # flat_index_document_filter_test.go:10-91
    f = _flat(O.L2, [[float(i), 0, 0] for i in range(10)])
    ids, _ = f.search([0, 0, 0], k=10, filter_ids=[2, 4, 6])
    assert sorted(ids.tolist()) == [2, 4, 6]
    ids, _ = f.search([0, 0, 0], k=10, filter_ids=[])          # empty filter == no filter
    assert len(ids) == 10
    ids, _ = f.search([0, 0, 0], k=10, filter_ids=[999])       # nothing eligible
    assert len(ids) == 0


def test_flat_soft_delete_and_flush():
    # flat_index_test.go:343-434
    f = _flat(O.L2, [[float(i), 0, 0] for i in range(5)])
    f.remove(2)
    ids, _ = f.search([0, 0, 0], k=10)
    assert 2 not in ids.tolist() and len(ids) == 4
    with pytest.raises(O.OracleError):
        f.remove(2)                 # already deleted
    with pytest.raises(O.OracleError):
        f.remove(77)                # not found
    assert len(f) == 5              # still stored until Flush
    f.flush()
    assert len(f) == 4
    ids, _ = f.search([0, 0, 0], k=10)
    assert ids.tolist() == [1, 3, 4, 5]


def test_flat_ties_keep_scan_order():
    # sort.Slice is an insertion sort (stable) for n <= 12: ties come back in insertion order
    f = _flat(O.L2SQ, [[1, 0], [0, 1], [-1, 0], [0, -1], [2, 0]])
    ids, sc = f.search([0, 0], k=5)
    assert ids.tolist() == [1, 2, 3, 4, 5]


def test_flat_multi_query_execute_semantics():
    # flat_index_search_test.go:229-278: Execute unions per-query top-k and aggregates by ID
    f = _flat(O.L2, [[1, 0, 0], [0, 1, 0], [0, 0, 1]])
    all_ids, all_sc = [], []
    for q in ([1, 0, 0], [0, 1, 0]):
        i, s = f.search(q, k=1)
        all_ids += i.tolist(); all_sc += s.tolist()
    ids, sc = O.aggregate("sum", all_ids, all_sc)
    k = O.sanitize_k(1, len(ids))
    assert k == 1 and len(set(ids.tolist())) == len(ids)


# ---- ivf_index_search_test.go:8-311 (shapes / sanity: what the reference pins) ------------
def test_ivf_search_basics():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((400, 8)).astype(np.float32)
    ivf = O.IVF(8, 10, O.L2)
    with pytest.raises(O.OracleError) as e:
        ivf.search(x[0], k=5)
    assert e.value.code == O.ERR_NOT_TRAINED          # "index must be trained before searching"
    with pytest.raises(O.OracleError):
        ivf.train(x[:5])                               # ivf_index.go:211 needs >= nlist vectors
    ivf.train(x)
    ivf.add(np.arange(1, 401), x.copy())
    assert ivf.default_nprobes() == 3                  # int(sqrt(10))
    ids, sc = ivf.search(x[17], k=5, nprobes=10)       # all lists == exhaustive
    flat = O.Flat(8, O.L2); flat.add(np.arange(1, 401), x.copy())
    fi, fs = flat.search(x[17], k=5)
    assert ids.tolist() == fi.tolist() and np.array_equal(sc, fs)
    assert ids[0] == 18 and sc[0] == 0.0
    ids2, _ = ivf.search(x[17], k=5, nprobes=0)        # <= 0 -> nlist (ivf_index_search.go:233-236)
    assert ids2.tolist() == ids.tolist()
    ids3, sc3 = ivf.search(x[17], k=5, nprobes=2)
    assert ids3[0] == 18 and all(sc3[i] <= sc3[i + 1] for i in range(len(sc3) - 1))


# ---- pq / ivfpq / hnsw: the reference pins counts, errors and "exact match ranks first" ---
def test_pq_basics():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((600, 16)).astype(np.float32)
    assert O.lib().co_pq_new(16, O.L2, 5, 8) is None   # dim % M != 0 (pq_index.go:144)
    assert O.lib().co_pq_new(16, O.L2, 4, 17) is None  # Nbits > 16 (pq_index.go:152)
    pq = O.PQ(16, O.L2, 4, 4)
    with pytest.raises(O.OracleError):
        pq.train(x[:3])                                # needs >= Ksub
    pq.train(x)
    pq.add(np.arange(1, 601), x.copy())
    assert pq.codes().shape == (600, 4) and pq.codes().max() < 16
    ids, sc = pq.search(x[5], k=10)
    assert len(ids) == 10 and all(sc[i] <= sc[i + 1] for i in range(9))
    # score is sqrt(sum of LUT entries) whatever the metric (pq_index_search.go:290-295)
    cb = pq.codebooks(); code = pq.codes()[ids[0] - 1]
    tot = np.float32(0)
    for m in range(4):
        sub = x[5, m * 4:(m + 1) * 4]; c = cb[m, code[m]]
        acc = np.float32(0)
        for j in range(4):
            dlt = np.float32(sub[j] - c[j]); acc = np.float32(acc + np.float32(dlt * dlt))
        tot = np.float32(tot + acc)
    assert sc[0] == np.float32(math.sqrt(float(tot)))


def test_ivfpq_basics():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((500, 16)).astype(np.float32)
    ix = O.IVFPQ(16, O.L2, 8, 4, 4)
    with pytest.raises(O.OracleError):
        ix.train(x[:79])                               # needs nlist*10 (ivfpq_index.go:185)
    ix.train(x)
    ix.add(np.arange(1, 501), x.copy())
    assert ix.total() == 500 and ix.default_nprobes() == 2
    ids, sc = ix.search(x[3], k=10, nprobes=8)
    assert len(ids) == 10 and all(sc[i] <= sc[i + 1] for i in range(9))
    assert O.IVFPQ.last_scanned() == 500


def test_hnsw_basics():
    rng = np.random.default_rng(4)
    x = rng.standard_normal((300, 8)).astype(np.float32)
    h = O.HNSW(8, O.L2, 8, 50, 50)
    assert len(h.search(x[0], k=5)[0]) == 0            # empty index -> empty result
    lv = O.hnsw_random_levels(300, 8, seed=11)
    h.add(np.arange(1, 301), x.copy(), lv)
    assert len(h) == 300 and h.entry_point == 1        # entry point is never promoted
    assert h.max_level == int(lv.max())
    ids, sc = h.search(x[0], k=5)
    assert ids[0] == 1 and sc[0] == 0.0                # the entry point itself is always reached
    assert all(sc[i] <= sc[i + 1] for i in range(len(sc) - 1))
    ids_all, _ = h.search(x[42], k=0, ef_search=300)   # k<=0 -> everything the beam kept
    assert len(ids_all) <= 300
    _, levels, _, layers = h.export()
    offs0, _ = layers[0]
    deg0 = np.diff(offs0)
    assert deg0.max() <= 16                            # layer-0 degree <= 2M
    ev, ex = O.HNSW.last_counters()
    assert ev > 0 and ex > 0


def test_hnsw_reference_accuracy_test():
    # hnsw_index_search_test.go:942-990 TestHNSWIndexSearchAccuracy
    h = O.HNSW(3, O.L2, 16, 200, 200)
    v = np.array([[1, 0, 0], [2, 0, 0], [0.5, 0, 0], [3, 0, 0]], np.float32)
    h.add([1, 2, 3, 4], v.copy(), [0, 0, 0, 0])
    ids, sc = h.search([0, 0, 0], k=4)
    assert len(ids) == 4 and ids[0] == 3
    assert sc.tolist() == [0.5, 1.0, 2.0, 3.0]


@pytest.mark.parametrize("seed", [0, 1])
def test_hnsw_reference_recall_test(seed):
    # hnsw_index_search_test.go:993-1040 TestHNSWIndexSearchRecall (levels from a seeded RNG)
    h = O.HNSW(10, O.L2, 16, 200, 200)
    n = 500
    x = np.array([[float((i * 10 + j) % 100) for j in range(10)] for i in range(n)], np.float32)
    h.add(np.arange(1, n + 1), x.copy(), O.hnsw_random_levels(n, 16, seed))
    q = np.array([float(j % 100) for j in range(10)], np.float32)
    ids, sc = h.search(q, k=10)
    assert len(ids) == 10 and all(s <= 500 for s in sc)


def test_hnsw_back_edge_drop_quirk():
    """hnsw_index.go:282-284 runs insertNode BEFORE idx.nodes[id] = node, so pruneConnections
    (hnsw_index.go:680-682, `idx.nodes[nid] == nil`) drops the node being inserted whenever a
    neighbour's list overflows.  With M=1 (layer-0 cap 2) the third neighbour of a node can never
    be linked back: node 1 keeps exactly the first two back-edges it ever received."""
    h = O.HNSW(2, O.L2SQ, 1, 10, 10)
    v = np.array([[0, 0], [1, 0], [0, 1], [0.1, 0.1], [0.05, 0.0]], np.float32)
    h.add([1, 2, 3, 4, 5], v.copy(), [0, 0, 0, 0, 0])
    _, _, _, layers = h.export()
    offs, nbrs = layers[0]
    assert nbrs[offs[0]:offs[1]].tolist() == [2, 3]    # later, closer nodes 4 and 5 were dropped


def test_hnsw_reference_efsearch_recall_test():
    # hnsw_index_search_test.go:1149-1207: query == first inserted vector (the entry point)
    h = O.HNSW(10, O.L2, 16, 200, 50)
    x = np.array([[float(i * 10 + j) for j in range(10)] for i in range(100)], np.float32)
    h.add(np.arange(1, 101), x.copy(), O.hnsw_random_levels(100, 16, 5))
    lo, _ = h.search(x[0], k=10, ef_search=10)
    hi, hs = h.search(x[0], k=10, ef_search=100)
    assert len(lo) > 0 and len(hi) > 0 and hi[0] == 1 and hs[0] == 0.0


def test_golden_reference_kats_file():
    """tests/golden/reference_kats.json (literals transcribed from the reference's tests) replayed on the oracle."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")
    with open(path) as f:
        kats = json.load(f)
    codes = {"l2": 0, "l2_squared": 1, "cosine": 2}
    for c in kats["distance"]:
        got = O.distance(codes[c["metric"]], np.asarray(c["a"], np.float32), np.asarray(c["b"], np.float32))
        assert abs(got - c["want"]) <= c["tol"], c
    for c in kats["normalize"]:
        if "error" in c:
            with pytest.raises(Exception):
                O.normalize(np.asarray(c["in"], np.float32))
        else:
            assert np.allclose(O.normalize(np.asarray(c["in"], np.float32)), c["want"], atol=c["tol"])
    for c in kats["norm"]:
        assert abs(O.norm(np.asarray(c["in"], np.float32)) - c["want"]) <= EPS, c
    for c in kats["normalize_helper"]:
        assert np.allclose(O.normalize(np.asarray(c["in"], np.float32)), c["want"], atol=EPS), c
    hd = kats["high_dimensional"]["dim"]
    a = (np.arange(hd) % 10).astype(np.float32)
    b = ((np.arange(hd) + 1) % 10).astype(np.float32)
    assert np.isfinite(O.distance(0, a, b)) and np.isfinite(O.distance(1, a, b))
    assert O.distance(1, a, b) == 692.0 * 1.0 + 76.0 * 81.0                       # 692 elements differ by 1, 76 by 9 (exact in fp32)
    assert np.isfinite(O.distance(2, O.normalize(a), O.normalize(b)))
    bc = kats["batch_consistency"]
    for qv in bc["queries"]:
        qv = np.asarray(qv, np.float32)
        assert O.distance(0, qv, np.asarray(bc["target"], np.float32)) == O.norm(qv)          # |q - 0| is the norm, same loop
        assert O.distance(1, qv, np.asarray(bc["target"], np.float32)) == float(np.float32(np.dot(qv, qv)))
    from comet_b200 import capi
    kc = kats["kind_constants"]
    assert capi.METRICS == {kc["l2"]: capi.L2, kc["l2_squared"]: capi.L2SQ, kc["cosine"]: capi.COSINE}
    for c in kats["sanitize_k"]:
        assert O.sanitize_k(c["k"], c["max"]) == c["want"], c
    for c in kats["flat_search"]:
        o = O.Flat(3, codes[c["metric"]])
        for id_, row in c["rows"].items():
            o.add([int(id_)], np.asarray([row], np.float32))
        ids, sc = o.search(np.asarray(c["query"], np.float32), k=c["k"])
        if "want_ids" in c:
            assert ids.tolist() == c["want_ids"] and sc.tolist() == c["want_scores"]
        else:
            assert ids[0] == c["want_first_id"] and sc[0] == c["want_first_score"]
