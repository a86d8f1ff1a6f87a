"""GPU parity tests of the tensor-core flat path (bf16 tcgen05 candidate pass + reference-order
re-score): results must be BIT-IDENTICAL to the CPU oracle, like the exact scan's."""
import os

import numpy as np
import pytest

from comet_b200 import capi
from oracle import oracle_py as O
from tests.parity import assert_same_results

pytestmark = pytest.mark.gpu


def make_pair(n, d, metric, seed, cta_group):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    os.environ["COMET_B200_CTA_GROUP"] = str(cta_group)
    try:
        g = capi.FlatIndex(d, metric)
    finally:
        os.environ.pop("COMET_B200_CTA_GROUP", None)
    g.add(ids, x.copy())
    o = O.Flat(d, metric)
    o.add(ids, x.copy())
    return g, o, rng


def check(g, o, q, k, **kw):
    ids, sc, cnt = g.search(q, k=k, path=capi.PATH_TENSOR, **kw)
    st = g.last_stats()
    assert st["path_used"] == capi.PATH_TENSOR
    O.set_threads(os.cpu_count() or 1)
    oi, os_, oc = o.search_batch(q, k, threshold=kw.get("threshold", 0.0))
    for i in range(q.shape[0]):
        assert_same_results(ids[i], sc[i], cnt[i], oi[i, :oc[i]], os_[i, :oc[i]], what=f"query {i}")
    return st


# cta_group 2 runs the query-resident candidate pass (queries in tensor memory, flat_gemm_ts.cu); 1 and "2s"
# (COMET_B200_NO_TS) the pass that stages both operands in shared memory (the one rows wider than 768 use)
@pytest.mark.parametrize("cta_group", [1, 2, "2s"])
@pytest.mark.parametrize("metric", [capi.L2SQ, capi.COSINE, capi.L2])
def test_tensor_path_matches_oracle(metric, cta_group, monkeypatch):
    if cta_group == "2s":
        monkeypatch.setenv("COMET_B200_NO_TS", "1")
        cta_group = 2
    g, o, rng = make_pair(33_000, 128, metric, 11 + metric, cta_group)
    q = rng.standard_normal((260, 128)).astype(np.float32)
    st = check(g, o, q, 10)
    assert st["fallback_queries"] == 0
    check(g, o, q[:70], 100)


@pytest.mark.parametrize("cta_group", [1, 2, "2s"])
def test_tensor_path_dim768_k100(cta_group, monkeypatch):
    if cta_group == "2s":
        monkeypatch.setenv("COMET_B200_NO_TS", "1")
        cta_group = 2
    g, o, rng = make_pair(24_000, 768, capi.COSINE, 5, cta_group)
    q = rng.standard_normal((300, 768)).astype(np.float32)
    st = check(g, o, q, 100)
    assert st["fallback_queries"] == 0


def test_tensor_path_odd_dim_and_ragged_tail():
    # dim not a multiple of 64 (bf16 shadow is zero padded), n not a multiple of the row tile
    g, o, rng = make_pair(20_000 + 77, 100, capi.L2SQ, 3, 2)
    q = rng.standard_normal((65, 100)).astype(np.float32)
    check(g, o, q, 7)


def test_tensor_path_with_deletes_and_threshold():
    g, o, rng = make_pair(30_000, 64, capi.L2, 9, 2)
    for i in range(1, 2000, 7):
        g.remove(i)
        o.remove(i)
    q = rng.standard_normal((128, 64)).astype(np.float32)
    check(g, o, q, 20)
    check(g, o, q, 50, threshold=9.5)


def test_tensor_path_ties_fall_back_to_exact():
    # the reference benchmark data (flat_index_document_filter_test.go:188-200): 100 distinct rows
    # repeated -> every candidate bound is met by thousands of rows; the overflow must be detected
    # and those queries answered by the exact scan, still bit-identical to the oracle.
    n, d = 20_000, 128
    x = np.repeat((np.arange(n) % 100).astype(np.float32)[:, None], d, axis=1)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    g = capi.FlatIndex(d, capi.L2)
    g.add(ids, x.copy())
    o = O.Flat(d, capi.L2)
    o.add(ids, x.copy())
    q = np.ones((64, d), np.float32)
    ids_g, sc_g, cnt_g = g.search(q, k=10, path=capi.PATH_TENSOR)
    oi, os_, oc = o.search_batch(q, 10)
    for i in range(len(q)):
        assert_same_results(ids_g[i], sc_g[i], cnt_g[i], oi[i, :oc[i]], os_[i, :oc[i]], what=f"query {i}")


def test_auto_path_policy():
    # CM_PATH_AUTO (measured in tools/nq_sweep.py): the candidate pass at every batch size once the index has 65,536
    # rows -- one query included -- the exact scan below that, for k > 256 and under filters that leave few rows
    g, o, rng = make_pair(66_000, 64, capi.COSINE, 21, 2)
    q = rng.standard_normal((64, 64)).astype(np.float32)
    for nq in (64, 8, 1):
        ids_g, sc_g, cnt_g = g.search(q[:nq], k=10)
        assert g.last_stats()["path_used"] == capi.PATH_TENSOR
        oi, os_, oc = o.search_batch(q[:nq], 10)
        for i in range(nq):
            assert_same_results(ids_g[i], sc_g[i], cnt_g[i], oi[i, :oc[i]], os_[i, :oc[i]], what=f"nq {nq} query {i}")
    g.search(q[:8], k=300)
    assert g.last_stats()["path_used"] == capi.PATH_EXACT
    small, _, _ = make_pair(20_000, 64, capi.COSINE, 22, 2)
    small.search(q, k=10)
    assert small.last_stats()["path_used"] == capi.PATH_EXACT


def test_document_filters_on_the_device():
    # WithDocumentIDs (flat_index_search.go:255-263): selective filters are looked up and only their rows are scored;
    # the others become a skip mask -- under the exact scan or, when they keep at least half of the rows, the tensor path
    g, o, rng = make_pair(70_000, 48, capi.L2, 23, 2)
    n = 70_000
    q = rng.standard_normal((9, 48)).astype(np.float32)
    ids_all = np.arange(1, n + 1, dtype=np.uint32)
    for dead in (10, 11, 50_000):
        g.remove(dead)
        o.remove(dead)

    def both(filt, k, **kw):
        gi, gs, gc = g.search(q, k=k, filter_ids=filt, **kw)
        for i in range(len(q)):
            oi, os_ = o.search(q[i], k=k, filter_ids=filt, **kw)
            assert_same_results(gi[i], gs[i], gc[i], oi, os_, what=f"query {i}")
        return g.last_stats()

    sel = rng.choice(ids_all, 300, replace=False)
    sel = np.concatenate([sel, sel[:20], np.array([10, 11, 999_999, 0], np.uint32)]).astype(np.uint32)   # repeats, deleted, unknown
    st = both(sel, 25)
    assert st["passes"] == 0                               # no pass over the corpus
    both(sel, 1000)                                        # k beyond the filtered rows: everything that passes, in order
    both(sel, 25, threshold=float(np.sort(((g.get_rows(np.arange(2000)) - q[0]) ** 2).sum(1))[5]) ** 0.5 + 3.0)
    both(np.array([10, 11], np.uint32), 5)                 # only deleted rows: nothing
    both(ids_all[::16].copy(), 40)                         # 1/16 of the rows: still the gather path
    st = both(ids_all[::3].copy(), 40)                     # a third: skip mask + exact scan
    assert st["path_used"] == capi.PATH_EXACT and st["passes"] > 0
    st = both(np.delete(ids_all, np.arange(0, n, 5)), 40)  # four fifths: skip mask folded into the candidate pass
    assert st["path_used"] == capi.PATH_TENSOR and st["fallback_queries"] == 0
    # node IDs may repeat (Add does not reject them): a filter ID selects every row that carries it
    x2 = rng.standard_normal((3, 48)).astype(np.float32)
    dup = np.array([sel[0], sel[0], sel[1]], np.uint32)
    g.add(dup, x2.copy())
    o.add(dup, x2.copy())
    both(sel, 25)
    g.flush()
    o.flush()
    both(sel, 25)


def test_tensor_path_select_shapes_and_candidate_count(monkeypatch):
    # The select kernel has two launch shapes (two CTAs per SM staging up to 12288 keys; opt-in: four CTAs per SM
    # staging up to 6144, chosen from the previous call's staged-key counts).  Both must give the oracle's answer,
    # call after call, and the device-side count of re-scored candidates must be at least K per query.
    g, o, rng = make_pair(40_000, 96, capi.COSINE, 31, 2)
    q = rng.standard_normal((130, 96)).astype(np.float32)
    for small in ("0", "1", "1"):
        monkeypatch.setenv("COMET_B200_SEL_SMALL", small)
        st = check(g, o, q, 25)
        assert st["fallback_queries"] == 0
        assert st["candidates"] >= 25 * len(q)
    monkeypatch.setenv("COMET_B200_NO_DENSE", "1")      # phase A through the sparse emission path
    check(g, o, q, 25)


def test_tensor_path_rows_wider_than_tensor_memory():
    # 800 dims: a query row no longer fits beside the accumulators in tensor memory -> both operands are staged
    g, o, rng = make_pair(17_000, 800, capi.L2SQ, 41, 2)
    q = rng.standard_normal((64, 800)).astype(np.float32)
    check(g, o, q, 10)


def test_tensor_path_many_query_blocks():
    # 1030 queries: two device chunks (1024 + 6), the first with four query blocks spread over the CTA pairs
    g, o, rng = make_pair(20_000, 64, capi.COSINE, 43, 2)
    q = rng.standard_normal((1030, 64)).astype(np.float32)
    check(g, o, q, 5)
