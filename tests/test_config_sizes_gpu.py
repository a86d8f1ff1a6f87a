"""Parity at the sizes BASELINE.json quotes its configurations on: every performance claim of
DESIGN.md has a green bit-exact comparison with the CPU oracle at its own size.

  C2  Flat 1M x 768, K = 100, ALL 512 queries of a batch through the tensor path, cosine and L2
      (flat_index_search.go:221-294);
  C2' a clustered corpus with near-duplicate rows through the tensor path: exactness, and the number of candidates
      the exact re-score saw per query / the queries that fell back to the exact scan are RECORDED;
  C3  IVFPQ, 768-d, nlist 4096, nprobes 32, M 96, nbits 8, K = 100: the device-trained centroids, codebooks and codes
      are exported into the oracle (its ReadFrom-style loader) and 16 queries compared (ivfpq_index_search.go:231-390);
  C4  HNSW 1M x 768, M 16, ef 128, K = 10 on an exact 32-NN graph loaded into both sides: ids, score bits and the
      work counters of 16 queries (hnsw_index_search.go:248-354, hnsw_index.go:565-629).

Sizes can be lowered for a quick local run: COMET_TEST_C2_N, COMET_TEST_C3_N, COMET_TEST_C4_N."""
import json
import os
import time

import numpy as np
import pytest

from comet_b200 import capi
from oracle import oracle_py as O
from tests.parity import assert_same_results

pytestmark = pytest.mark.gpu

D = 768
C2_N = int(os.environ.get("COMET_TEST_C2_N", 1_000_000))
C3_N = int(os.environ.get("COMET_TEST_C3_N", 2_500_000))
C4_N = int(os.environ.get("COMET_TEST_C4_N", 1_000_000))
RECORD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def record(name, obj):
    """Keep what the run measured next to the other GPU artefacts (copied into profiles/ by hand)."""
    if os.path.isdir(RECORD):
        with open(os.path.join(RECORD, name), "w") as f:
            json.dump(obj, f, indent=1)


@pytest.fixture(scope="module")
def corpus():
    rng = np.random.default_rng(20261017)
    return rng.standard_normal((C2_N, D), dtype=np.float32)


@pytest.mark.parametrize("metric", [capi.COSINE, capi.L2], ids=["cosine", "l2"])
def test_c2_tensor_path_all_512_queries(corpus, metric):
    ids = np.arange(1, C2_N + 1, dtype=np.uint32)
    g = capi.FlatIndex(D, metric)
    g.add(ids, corpus, writeback=False)
    x = corpus.copy()
    o = O.Flat(D, metric)
    o.add(ids, x)                                   # normalises its copy in place, like FlatIndex.Add
    q = np.random.default_rng(7).standard_normal((512, D), dtype=np.float32)
    gi, gs, gc = g.search(q, k=100, path=capi.PATH_TENSOR)
    st = g.last_stats()
    assert st["path_used"] == capi.PATH_TENSOR and st["fallback_queries"] == 0
    O.set_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    oi, os_, oc = o.search_batch(q, 100)
    dt = time.perf_counter() - t0
    for i in range(512):
        assert_same_results(gi[i], gs[i], gc[i], oi[i, :oc[i]], os_[i, :oc[i]], what=f"query {i}")
    record(f"r02_parity_c2_{'cosine' if metric == capi.COSINE else 'l2'}.json",
           {"rows": C2_N, "dim": D, "k": 100, "queries_checked": 512, "ids_bit_exact": True, "scores_bit_exact": True,
            "rescored_candidates": int(st["candidates"]), "fallback_queries": int(st["fallback_queries"]),
            "oracle_seconds": dt, "oracle_threads": os.cpu_count()})


def test_c2_clustered_near_duplicate_corpus_is_exact_and_recorded():
    # rows on a 32-dimensional manifold (one cloud, strongly correlated scores) with every fifth row a near-duplicate
    # of another one: far more rows sit inside the candidate band than on i.i.d. Gaussians
    n, nq = min(C2_N, 262_144), 256
    rng = np.random.default_rng(5)
    W = rng.standard_normal((32, D), dtype=np.float32)
    x = rng.standard_normal((n, 32), dtype=np.float32) @ W + 0.05 * rng.standard_normal((n, D), dtype=np.float32)
    dup = np.arange(0, n, 5)
    x[dup] = x[(dup * 7 + 3) % n] + 1e-3 * rng.standard_normal((len(dup), D), dtype=np.float32)
    q = rng.standard_normal((nq, 32), dtype=np.float32) @ W
    ids = np.arange(1, n + 1, dtype=np.uint32)
    out = {}
    for metric, name in ((capi.L2, "l2"), (capi.COSINE, "cosine")):
        g = capi.FlatIndex(D, metric)
        g.add(ids, x, writeback=False)
        o = O.Flat(D, metric)
        o.add(ids, x.copy())
        gi, gs, gc = g.search(q, k=100, path=capi.PATH_TENSOR)
        st = g.last_stats()
        O.set_threads(os.cpu_count() or 1)
        oi, os_, oc = o.search_batch(q, 100)
        for i in range(nq):
            assert_same_results(gi[i], gs[i], gc[i], oi[i, :oc[i]], os_[i, :oc[i]], what=f"{name} query {i}")
        out[name] = {"rows": n, "queries": nq, "k": 100, "rescored_candidates_per_query": st["candidates"] / nq,
                     "fallback_queries": int(st["fallback_queries"])}
    record("r02_clustered_candidates.json", out)


def test_c3_ivfpq_device_state_in_the_oracle():
    n, nlist, nprobes, M = C3_N, 4096, 32, 96
    rng = np.random.default_rng(11)
    W = rng.standard_normal((32, D), dtype=np.float32)

    def rows(m):
        return rng.standard_normal((m, 32), dtype=np.float32) @ W + 0.05 * rng.standard_normal((m, D), dtype=np.float32)

    ix = capi.IVFPQIndex(D, capi.L2, nlist, M, 8)
    ix.train(rows(nlist * 16))
    lists = np.zeros(n, np.int32)
    slab = 500_000
    for s0 in range(0, n, slab):
        m = min(slab, n - s0)
        lists[s0:s0 + m] = ix.add(np.arange(s0 + 1, s0 + m + 1, dtype=np.uint32), rows(m), writeback=False)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    o = O.IVFPQ(D, capi.L2, nlist, M, 8)
    o.set_trained(*ix.trained_state())
    o.load_codes(ids, ix.codes(), lists)            # arrival order: every list gets its members in stored order
    q = rows(16)
    gi, gs, gc = ix.search(q, k=100, nprobes=nprobes)
    scanned = []
    for i in range(16):
        oi, os_ = o.search(q[i], k=100, nprobes=nprobes)
        scanned.append(int(O.IVFPQ.last_scanned()))
        assert_same_results(gi[i], gs[i], gc[i], oi, os_, what=f"query {i}")
    record("r02_parity_c3_ivfpq.json", {"rows": n, "dim": D, "nlist": nlist, "nprobes": nprobes, "M": M, "nbits": 8, "k": 100,
                                        "queries_checked": 16, "ids_bit_exact": True, "scores_bit_exact": True,
                                        "codes_scanned_per_query": float(np.mean(scanned))})


def test_c4_hnsw_same_graph_same_walk():
    n = C4_N
    rng = np.random.default_rng(13)
    W = rng.standard_normal((24, D), dtype=np.float32)
    x = rng.standard_normal((n, 24), dtype=np.float32) @ W
    q = rng.standard_normal((16, 24), dtype=np.float32) @ W
    ids = np.arange(1, n + 1, dtype=np.uint32)
    flat = capi.FlatIndex(D, capi.L2)
    flat.add(ids, x, writeback=False)
    nbr = np.zeros((n, 32), np.uint32)
    for s0 in range(0, n, 16384):
        gi, _, _ = flat.search(x[s0:s0 + 16384], k=33)
        nbr[s0:s0 + 16384] = gi[:, 1:33]
    del flat
    levels = np.zeros(n, np.int32)
    off = np.arange(n + 1, dtype=np.int64) * 32
    g = capi.HNSWIndex(D, capi.L2, 16, 100, 128)
    g.load_graph(ids, x, levels, [(off, nbr.ravel())], 1, 0)
    o = O.HNSW(D, capi.L2, 16, 100, 128)
    o.load_graph(ids, x, levels, off, nbr.ravel(), 1, 0)
    gi, gs, gc, work = g.search(q, k=10, ef_search=128, with_work=True)
    evals = []
    for i in range(16):
        oi, os_ = o.search(q[i], k=10, ef_search=128)
        ev, ex = O.HNSW.last_counters()
        assert_same_results(gi[i], gs[i], gc[i], oi, os_, what=f"query {i}")
        assert (int(work[i, 0]), int(work[i, 1])) == (ev, ex), f"query {i}: work counters {work[i]} != oracle {(ev, ex)}"
        evals.append(ev)
    record("r02_parity_c4_hnsw.json", {"rows": n, "dim": D, "degree": 32, "ef": 128, "k": 10, "queries_checked": 16,
                                       "ids_bit_exact": True, "scores_bit_exact": True, "work_counters_equal": True,
                                       "distance_evaluations_per_query": float(np.mean(evals))})
