"""GPU parity tests of the PQ and IVFPQ paths against the CPU oracle on a SHARED trained index (the oracle
trains codebooks / centroids with the reference's deterministic k-means; the device loads them): codes,
list assignment, ids, ranks and score bits must be identical.  Ties are frequent in ADC scores (few
distinct table sums) -- both sides order them by candidate number."""
import numpy as np
import pytest

from comet_b200 import capi
from oracle import oracle_py as O
from tests.parity import assert_same_results

pytestmark = pytest.mark.gpu


def data(n, d, seed, shift=0.0):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal((n, d)).astype(np.float32) + np.float32(shift)), rng


def pq_pair(n, d, metric, M, nbits, seed):
    x, rng = data(n, d, seed, 0.2 if metric == capi.COSINE else 0.0)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    o = O.PQ(d, metric, M, nbits)
    o.train(x[:max(1 << nbits, 600)].copy())
    g = capi.PQIndex(d, metric, M, nbits)
    g.set_codebooks(o.codebooks())
    o.add(ids, x.copy())
    g.add(ids, x.copy())
    return g, o, rng


def check_pq(g, o, q, k, **kw):
    ids, sc, cnt = g.search(q, k=k, **kw)
    for i in range(q.shape[0]):
        oi, os_ = o.search(q[i], k=k, **kw)
        assert_same_results(ids[i], sc[i], cnt[i], oi, os_, what=f"query {i}")


@pytest.mark.parametrize("metric", [capi.L2, capi.COSINE])
def test_pq_codes_and_search_match_oracle(metric):
    g, o, rng = pq_pair(9000, 32, metric, 8, 6, 40 + metric)
    assert np.array_equal(g.codes(), o.codes())
    q = rng.standard_normal((13, 32)).astype(np.float32)
    check_pq(g, o, q, 10)
    check_pq(g, o, q[:4], 300)
    check_pq(g, o, q[:3], 25, threshold=4.0 if metric == capi.L2 else 0.9)
    check_pq(g, o, q[:3], 25, filter_ids=np.arange(5, 9000, 7, dtype=np.uint32))


def test_pq_m96_nbits8_dim768():
    g, o, rng = pq_pair(3000, 768, capi.L2SQ, 96, 8, 5)
    assert np.array_equal(g.codes(), o.codes())
    q = rng.standard_normal((6, 768)).astype(np.float32)
    check_pq(g, o, q, 100)


def test_pq_delete_flush_and_errors():
    g, o, rng = pq_pair(5000, 16, capi.L2, 4, 4, 77)
    q = rng.standard_normal((5, 16)).astype(np.float32)
    for i in range(3, 1500, 4):
        g.remove(i)
        o.remove(i)
    check_pq(g, o, q, 20)
    g.flush()
    o.flush()
    assert len(g) == len(o)
    assert np.array_equal(g.codes(), o.codes())
    check_pq(g, o, q, 20)
    h = capi.PQIndex(16, capi.L2, 4, 4)
    with pytest.raises(capi.CometError) as e:
        h.search(q, k=1)
    assert e.value.code == capi.ERR_NOT_TRAINED and "index not trained" in e.value.msg
    with pytest.raises(capi.CometError):
        capi.PQIndex(10, capi.L2, 4, 8)            # dim not divisible by M (pq_index.go:146)
    with pytest.raises(capi.CometError):
        capi.PQIndex(16, capi.L2, 4, 256)          # the reference's own benchmark argument: rejected (pq_index.go:152)
    h.set_codebooks(o.codebooks())
    assert h.search(q, k=3)[2].tolist() == [0] * 5   # trained, empty


def ivfpq_pair(n, d, metric, nlist, M, nbits, seed):
    x, rng = data(n, d, seed, 0.2 if metric == capi.COSINE else 0.0)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    o = O.IVFPQ(d, metric, nlist, M, nbits)
    o.train(x[:max(nlist * 10, (1 << nbits) * 2, 800)].copy())
    g = capi.IVFPQIndex(d, metric, nlist, M, nbits)
    g.set_trained(o.centroids(), o.codebooks())
    o.add(ids, x.copy())
    lists = g.add(ids, x.copy())
    return g, o, rng, lists


def check_ivfpq(g, o, q, k, nprobes, **kw):
    ids, sc, cnt = g.search(q, k=k, nprobes=nprobes, **kw)
    for i in range(q.shape[0]):
        oi, os_ = o.search(q[i], k=k, nprobes=nprobes, **kw)
        assert_same_results(ids[i], sc[i], cnt[i], oi, os_, what=f"query {i}")


@pytest.mark.parametrize("metric", [capi.L2, capi.COSINE])
def test_ivfpq_lists_codes_and_search_match_oracle(metric):
    g, o, rng, lists = ivfpq_pair(8000, 32, metric, 20, 8, 5, 60 + metric)
    ol = o.lists()
    codes = g.codes()
    for l in range(20):
        mine = np.nonzero(lists == l)[0]
        assert np.array_equal((mine + 1).astype(np.uint32), ol[l][0]), f"list {l} membership"
        assert np.array_equal(codes[mine], ol[l][1]), f"list {l} residual codes"
    q = rng.standard_normal((11, 32)).astype(np.float32)
    for nprobes in (1, 4, 20):
        check_ivfpq(g, o, q, 10, nprobes)
    check_ivfpq(g, o, q[:3], 0, 2)
    check_ivfpq(g, o, q[:3], 50, 4, threshold=5.0 if metric == capi.L2 else 1.1)
    check_ivfpq(g, o, q[:3], 50, 4, filter_ids=np.arange(2, 8000, 5, dtype=np.uint32))


def test_ivfpq_m96_dim768_and_delete_flush():
    g, o, rng, _ = ivfpq_pair(4000, 768, capi.L2SQ, 16, 96, 8, 9)
    assert g.default_nprobes() == o.default_nprobes() == 4
    q = rng.standard_normal((5, 768)).astype(np.float32)
    check_ivfpq(g, o, q, 100, 4)
    for i in range(1, 800, 3):
        g.remove(i)
        o.remove(i)
    check_ivfpq(g, o, q, 100, 4)
    g.flush()
    o.flush()
    assert len(g) == o.total()
    check_ivfpq(g, o, q, 100, 16)


# The ADC kernel has one variant per code-row width (M / 16 sixteen-byte words, each with its own number
# of rows per thread per round) plus a byte-wise generic one; a pair's codes are cut into slices when a
# launch would otherwise have too few CTAs.  Every variant, several rounds, compaction and slicing:
@pytest.mark.parametrize("d,M,nbits", [(64, 16, 8), (64, 32, 8), (128, 64, 8), (256, 128, 8), (96, 48, 8), (64, 16, 6)])
def test_pq_every_code_width_sliced_scan(d, M, nbits):
    g, o, rng = pq_pair(21000, d, capi.L2, M, nbits, 300 + M + nbits)
    assert np.array_equal(g.codes(), o.codes())
    q = rng.standard_normal((3, d)).astype(np.float32)
    check_pq(g, o, q, 100)
    check_pq(g, o, q[:2], 700)          # k above one round's appends: bigger candidate buffer
    check_pq(g, o, q[:1], 40, filter_ids=np.arange(3, 21000, 11, dtype=np.uint32))


def test_pq_m96_many_rounds_one_table_per_slice():
    g, o, rng = pq_pair(30000, 768, capi.L2, 96, 8, 11)
    q = rng.standard_normal((4, 768)).astype(np.float32)
    check_pq(g, o, q, 100)


def test_ivfpq_sliced_lists(monkeypatch):
    g, o, rng, _ = ivfpq_pair(9000, 64, capi.L2, 2, 16, 8, 123)
    q = rng.standard_normal((5, 64)).astype(np.float32)
    for slices in ("1", "3", "7"):
        monkeypatch.setenv("COMET_B200_ADC_SLICES", slices)
        check_ivfpq(g, o, q, 60, 2)
        check_ivfpq(g, o, q[:2], 0, 1)
    monkeypatch.setenv("COMET_B200_ADC_GENERIC", "1")
    check_ivfpq(g, o, q, 60, 2)


def test_adc_selection_under_heavy_ties():
    # A handful of distinct vectors repeated thousands of times: every ADC score is shared by thousands of codes, the
    # radix-select compaction cannot separate them (it must report that and let the sort decide), and the answer is
    # still the oracle's (score, candidate number) order.
    rng = np.random.default_rng(77)
    d, M, nbits = 64, 16, 8
    base = rng.standard_normal((5, d)).astype(np.float32)
    x = base[rng.integers(0, 5, 12_000)]
    x[::97] += rng.standard_normal((len(x[::97]), d)).astype(np.float32) * 0.01
    ids = np.arange(1, len(x) + 1, dtype=np.uint32)
    train = rng.standard_normal((600, d)).astype(np.float32)
    o = O.PQ(d, capi.L2, M, nbits)
    o.train(train.copy())
    g = capi.PQIndex(d, capi.L2, M, nbits)
    g.set_codebooks(o.codebooks())
    o.add(ids, x.copy())
    g.add(ids, x.copy())
    q = np.concatenate([base[:2], rng.standard_normal((2, d)).astype(np.float32)])
    check_pq(g, o, q, 50)
    check_pq(g, o, q, 300)
    oi = O.IVFPQ(d, capi.L2, 4, M, nbits)
    oi.train(np.concatenate([train, x[:400]]).copy())
    gi = capi.IVFPQIndex(d, capi.L2, 4, M, nbits)
    gi.set_trained(oi.centroids(), oi.codebooks())
    oi.add(ids, x.copy())
    gi.add(ids, x.copy())
    check_ivfpq(gi, oi, q, 50, 2)
    check_ivfpq(gi, oi, q[:2], 3000, 1)


def test_pq_and_ivfpq_k_all_beyond_the_per_cta_selection():
    # WithK(0) on more codes than a CTA can select from (k > 8192): sliced scan returning every key + a sort per query;
    # ADC scores tie massively (M = 4, 16 codewords), so this also checks tie order = candidate number
    g, o, rng = pq_pair(20000, 16, capi.L2, 4, 4, 71)
    q = rng.standard_normal((3, 16)).astype(np.float32)
    check_pq(g, o, q, 0)
    check_pq(g, o, q[:2], 15000)
    check_pq(g, o, q[:1], 0, threshold=12.0)
    g, o, rng, lists = ivfpq_pair(26000, 16, capi.L2, 3, 4, 4, 72)
    q = rng.standard_normal((3, 16)).astype(np.float32)
    check_ivfpq(g, o, q, 0, 3)
    check_ivfpq(g, o, q[:2], 12000, 2)


# The IVFPQ scan has a second form for 8-bit codes with M = 32 * PER ("ring": a lane owns PER sub-quantisers, the rows
# flow through the lanes; codes stored pre-skewed per list).  Every instantiated (PER, dsub / 4) pair, short / ragged /
# empty lists, lists long enough for several rounds and compactions, slices, threshold, filter, k = all -- against
# the oracle, and against the row-per-lane form on the same index (COMET_B200_ADC_RING=0 rebuilds the by-list copy).
@pytest.mark.parametrize("d,M", [(256, 32), (768, 32), (256, 64), (512, 64), (768, 64), (384, 96), (768, 96), (512, 128),
                                 (1024, 128)])
def test_ivfpq_ring_scan_every_variant(d, M, monkeypatch):
    g, o, rng, lists = ivfpq_pair(7000, d, capi.L2, 5, M, 8, 300 + M + d)
    sizes = np.bincount(lists, minlength=5)
    assert sizes.max() > 1300            # more than one buffer of candidates per pair: compaction runs
    q = rng.standard_normal((6, d)).astype(np.float32)
    check_ivfpq(g, o, q, 100, 5)
    check_ivfpq(g, o, q[:3], 10, 2)
    check_ivfpq(g, o, q[:2], 0, 3)
    check_ivfpq(g, o, q[:2], 500, 5, threshold=float(np.sqrt(2 * d) * 0.97))
    check_ivfpq(g, o, q[:2], 64, 5, filter_ids=np.arange(3, 7000, 7, dtype=np.uint32))
    for slices in ("2", "5"):
        monkeypatch.setenv("COMET_B200_ADC_SLICES", slices)
        check_ivfpq(g, o, q[:3], 100, 5)
    monkeypatch.delenv("COMET_B200_ADC_SLICES")
    ring = g.search(q, k=100, nprobes=5)
    monkeypatch.setenv("COMET_B200_ADC_RING", "0")
    rows = g.search(q, k=100, nprobes=5)
    monkeypatch.delenv("COMET_B200_ADC_RING")
    for a, b in zip(ring, rows):
        assert np.array_equal(a, b)
    again = g.search(q, k=100, nprobes=5)
    for a, b in zip(ring, again):
        assert np.array_equal(a, b)


def test_ivfpq_ring_scan_ragged_lists_and_updates():
    # tiny and empty lists (fewer rows than the 32-lane pipeline is deep), delete / flush / add between searches
    d, M, nlist = 768, 96, 40
    g, o, rng, lists = ivfpq_pair(900, d, capi.COSINE, nlist, M, 8, 41)
    sizes = np.bincount(lists, minlength=nlist)
    assert sizes.min() < 31
    q = rng.standard_normal((9, d)).astype(np.float32) + np.float32(0.2)
    check_ivfpq(g, o, q, 100, 40)
    check_ivfpq(g, o, q, 7, 1)
    for i in range(2, 900, 4):
        g.remove(i)
        o.remove(i)
    check_ivfpq(g, o, q, 100, 12)
    g.flush()
    o.flush()
    check_ivfpq(g, o, q, 100, 12)
    x2 = rng.standard_normal((333, d)).astype(np.float32) + np.float32(0.2)
    ids2 = np.arange(5000, 5333, dtype=np.uint32)
    g.add(ids2, x2.copy())
    o.add(ids2, x2.copy())
    check_ivfpq(g, o, q, 0, 40)


def test_ivfpq_ring_scan_k_all_beyond_the_per_cta_selection():
    g, o, rng, lists = ivfpq_pair(30000, 256, capi.L2, 2, 64, 8, 73)
    q = rng.standard_normal((2, 256)).astype(np.float32)
    check_ivfpq(g, o, q, 0, 2)
    check_ivfpq(g, o, q[:1], 12000, 1)


@pytest.mark.parametrize("d,M", [(256, 32), (512, 64), (768, 96), (1024, 128)])
def test_pq_ring_scan(d, M, monkeypatch):
    # the plain PQ index keeps a pre-skewed copy of its store for the same scan: slices of one long list, threshold,
    # filter, delete / flush / add in between, and the row-per-lane form on the same index
    g, o, rng = pq_pair(21000, d, capi.L2, M, 8, 500 + M)
    q = rng.standard_normal((5, d)).astype(np.float32)
    check_pq(g, o, q, 100)
    check_pq(g, o, q[:2], 0)
    check_pq(g, o, q[:2], 300, threshold=float(np.sqrt(2 * d) * 0.97))
    check_pq(g, o, q[:2], 64, filter_ids=np.arange(3, 21000, 7, dtype=np.uint32))
    for slices in ("1", "3"):
        monkeypatch.setenv("COMET_B200_ADC_SLICES", slices)
        check_pq(g, o, q[:3], 100)
    monkeypatch.delenv("COMET_B200_ADC_SLICES")
    ring = g.search(q, k=100)
    monkeypatch.setenv("COMET_B200_ADC_RING", "0")
    rows = g.search(q, k=100)
    monkeypatch.delenv("COMET_B200_ADC_RING")
    for a, b in zip(ring, rows):
        assert np.array_equal(a, b)
    for i in range(2, 6000, 5):
        g.remove(i)
        o.remove(i)
    check_pq(g, o, q, 100)
    g.flush()
    o.flush()
    x2 = rng.standard_normal((700, d)).astype(np.float32)
    ids2 = np.arange(50000, 50700, dtype=np.uint32)
    g.add(ids2, x2.copy())
    o.add(ids2, x2.copy())
    check_pq(g, o, q, 100)
