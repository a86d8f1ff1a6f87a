"""GPU parity tests of the flat path: CUDA (through the C ABI) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

from comet_b200 import capi
from oracle import oracle_py as O
from tests.parity import assert_same_results, bits

pytestmark = pytest.mark.gpu

METRICS = [capi.L2, capi.L2SQ, capi.COSINE]


def build_pair(x, metric, ids=None):
    n, d = x.shape
    ids = np.arange(1, n + 1, dtype=np.uint32) if ids is None else ids
    g = capi.FlatIndex(d, metric)
    g.add(ids, x.copy())
    o = O.Flat(d, metric)
    o.add(ids, x.copy())
    return g, o


def check_queries(g, o, q, k, path=capi.PATH_EXACT, **kw):
    ids, sc, cnt = g.search(q, k=k, path=path, **kw)
    for i in range(q.shape[0]):
        oi, os_ = o.search(q[i], k=k, **kw)
        assert_same_results(ids[i], sc[i], cnt[i], oi, os_, what=f"query {i}")


@pytest.mark.parametrize("metric", METRICS)
@pytest.mark.parametrize("dim", [1, 3, 32, 100, 128, 768])
def test_distance_pairs_bit_exact(metric, dim):
    rng = np.random.default_rng(dim * 7 + metric)
    a = rng.standard_normal((257, dim)).astype(np.float32)
    b = rng.standard_normal((257, dim)).astype(np.float32)
    if metric == capi.COSINE:
        a /= np.linalg.norm(a, axis=1, keepdims=True); b /= np.linalg.norm(b, axis=1, keepdims=True)
    got = capi.distance_pairs(metric, a, b)
    want = np.array([O.distance(metric, a[i], b[i]) for i in range(len(a))], np.float32)
    assert np.array_equal(bits(got), bits(want))


def test_distance_kats_on_device():
    # distance_test.go:87-145, 214-266, 335-387 replayed on the device kernels
    assert capi.distance_pairs(capi.L2, [0, 0], [3, 4])[0] == 5.0
    assert abs(capi.distance_pairs(capi.L2, [-1, -2], [1, 2])[0] - 4.472136) < 1e-6
    assert capi.distance_pairs(capi.L2SQ, [0, 0], [3, 4])[0] == 25.0
    assert capi.distance_pairs(capi.L2SQ, [-1, -2], [1, 2])[0] == 20.0
    assert capi.distance_pairs(capi.COSINE, [1, 0], [-1, 0])[0] == 2.0
    assert abs(capi.distance_pairs(capi.COSINE, [0.707107, 0.707107], [1, 0])[0] - 0.292893) < 1e-6


def test_preprocess_rows_bit_exact_and_zero_vector():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((100, 96)).astype(np.float32)
    want = np.stack([O.normalize(r) for r in x])
    got = capi.preprocess_rows(capi.COSINE, x.copy())
    assert np.array_equal(bits(got), bits(want))
    y = x.copy(); y[7] = 0
    with pytest.raises(capi.CometError) as e:
        capi.preprocess_rows(capi.COSINE, y)
    assert e.value.code == capi.ERR_ZERO_VECTOR
    assert np.array_equal(bits(y[:7]), bits(want[:7]))      # rows before the zero vector were normalised
    z = x.copy()
    assert np.array_equal(capi.preprocess_rows(capi.L2, z), x)   # no-op for l2


def test_reference_small_cases():
    # flat_index_search_test.go:10-48, 51-86, 348-389, 490-536 through the CUDA path
    g = capi.FlatIndex(3, capi.L2)
    g.add([1, 2, 3, 4], np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 0]], np.float32))
    ids, sc, cnt = g.search([1, 0, 0], k=2)
    assert cnt[0] == 2 and ids[0, 0] == 1 and sc[0, 0] == 0.0
    assert np.array_equal(g.get_vector(1), [1, 0, 0])
    g = capi.FlatIndex(3, capi.L2)
    g.add([1, 2, 3, 4], np.array([[1, 0, 0], [2, 0, 0], [4, 0, 0], [10, 0, 0]], np.float32))
    ids, sc, cnt = g.search([1, 0, 0], k=10, threshold=2.0)
    assert cnt[0] == 2
    g = capi.FlatIndex(3, capi.L2)
    g.add(np.arange(1, 6), np.array([[float(i), 0, 0] for i in range(5)], np.float32))
    for k, want in [(0, 5), (-1, 5), (3, 3), (5, 5), (100, 5), (1, 1)]:
        assert g.search([0, 0, 0], k=k)[2][0] == want
    g = capi.FlatIndex(3, capi.L2)
    g.add([1, 2, 3, 4], np.array([[5, 0, 0], [1, 0, 0], [10, 0, 0], [3, 0, 0]], np.float32))
    ids, sc, cnt = g.search([0, 0, 0], k=4)
    assert sc[0].tolist() == [1.0, 3.0, 5.0, 10.0] and ids[0].tolist() == [2, 4, 1, 3]


def test_errors_match_reference():
    g = capi.FlatIndex(4, capi.COSINE)
    with pytest.raises(capi.CometError) as e:
        g.add([1], np.zeros((1, 4), np.float32))
    assert e.value.code == capi.ERR_ZERO_VECTOR
    g.add([1, 2], np.array([[1, 0, 0, 0], [0, 2, 0, 0]], np.float32))
    with pytest.raises(capi.CometError) as e:
        g.search(np.zeros((1, 5), np.float32), k=1)
    assert e.value.code == capi.ERR_DIM_MISMATCH and "query dimension mismatch: expected 4, got 5" in e.value.msg
    with pytest.raises(capi.CometError) as e:
        g.search(np.zeros((1, 4), np.float32), k=1)           # zero query under cosine
    assert e.value.code == capi.ERR_ZERO_VECTOR
    with pytest.raises(capi.CometError):
        g.remove(99)
    with pytest.raises(capi.CometError):
        capi.FlatIndex(0, capi.L2)
    empty = capi.FlatIndex(4, capi.L2)
    assert empty.search(np.ones((2, 4), np.float32), k=3)[2].tolist() == [0, 0]


def test_cosine_add_normalises_callers_buffer():
    # flat_index.go:169-189 + SURVEY F7
    rows = np.array([[3, 4], [0, 5]], np.float32)
    g = capi.FlatIndex(2, capi.COSINE)
    g.add([1, 2], rows)
    want = np.stack([O.normalize([3, 4]), O.normalize([0, 5])])
    assert np.array_equal(bits(rows), bits(want))
    assert np.array_equal(bits(g.get_rows([0, 1])), bits(want))


@pytest.mark.parametrize("metric", METRICS)
def test_config1_10k_x_128(metric):
    # BASELINE configs[0]: Flat, 10K x 128, K=10, single query
    rng = np.random.default_rng(1)
    x = rng.standard_normal((10000, 128)).astype(np.float32)
    g, o = build_pair(x, metric)
    q = rng.standard_normal((1, 128)).astype(np.float32)
    check_queries(g, o, q, 10)
    q = rng.standard_normal((19, 128)).astype(np.float32)     # ragged batch: 8 + 8 + 3
    check_queries(g, o, q, 10)


def test_reference_benchmark_dataset_is_tie_saturated():
    # flat_index_document_filter_test.go:184-239: vec[j] = float32(i % 100), query all 1.0, K=10, Euclidean
    n, d = 10000, 128
    x = np.repeat((np.arange(n) % 100).astype(np.float32)[:, None], d, axis=1)
    g, o = build_pair(x, capi.L2)
    q = np.ones((1, d), np.float32)
    check_queries(g, o, q, 10)
    check_queries(g, o, q, 250)      # spans several 100-way tie groups


@pytest.mark.parametrize("dim", [3, 33, 100, 130])
@pytest.mark.parametrize("n", [1, 127, 128, 129, 1000])
def test_ragged_shapes(dim, n):
    rng = np.random.default_rng(dim * 1000 + n)
    x = rng.standard_normal((n, dim)).astype(np.float32)
    g, o = build_pair(x, capi.L2SQ)
    q = rng.standard_normal((5, dim)).astype(np.float32)
    check_queries(g, o, q, 7)
    check_queries(g, o, q, 0)        # WithK(0) -> everything


@pytest.mark.parametrize("metric", [capi.L2, capi.COSINE])
def test_small_table_many_query_groups(metric):
    # the shape of an IVF / IVFPQ coarse step: a table of a few dozen 128-row tiles scanned for hundreds of queries in ONE
    # launch; a query group then gets only a few CTAs (one wave for the launch), each walking several tiles
    rng = np.random.default_rng(4096 + metric)
    x = rng.standard_normal((3001, 48)).astype(np.float32) + np.float32(0.1)
    x[1500:1530] = x[10]                      # ties across tiles
    g, o = build_pair(x, metric)
    q = rng.standard_normal((603, 48)).astype(np.float32) + np.float32(0.1)
    ids, sc, cnt = g.search(q, k=32, path=capi.PATH_EXACT)
    for i in range(0, 603, 7):
        oi, os_ = o.search(q[i], k=32)
        assert_same_results(ids[i], sc[i], cnt[i], oi, os_, what=f"query {i}")


@pytest.mark.parametrize("k", [1, 100, 1000, 4096])
def test_k_sweep(k):
    rng = np.random.default_rng(k)
    x = rng.standard_normal((20000, 64)).astype(np.float32)
    g, o = build_pair(x, capi.L2)
    q = rng.standard_normal((3, 64)).astype(np.float32)
    check_queries(g, o, q, k)


def test_k_all_big():
    rng = np.random.default_rng(11)
    x = rng.standard_normal((6000, 48)).astype(np.float32)
    g, o = build_pair(x, capi.COSINE)
    q = rng.standard_normal((2, 48)).astype(np.float32)
    check_queries(g, o, q, 0)                       # 6000 > 4096: radix-sort path
    check_queries(g, o, q, 5000, threshold=1.0)


def test_threshold_filter_delete_flush():
    rng = np.random.default_rng(12)
    x = rng.standard_normal((5000, 40)).astype(np.float32)
    ids = rng.permutation(np.arange(100, 100 + 5000)).astype(np.uint32)
    g, o = build_pair(x, capi.L2, ids)
    q = rng.standard_normal((4, 40)).astype(np.float32)
    check_queries(g, o, q, 50, threshold=8.0)
    filt = ids[rng.choice(5000, 300, replace=False)]
    check_queries(g, o, q, 50, filter_ids=filt)
    check_queries(g, o, q, 50, filter_ids=[1, 2, 3])            # nothing eligible
    for victim in ids[:40]:
        g.remove(int(victim)); o.remove(int(victim))
    check_queries(g, o, q, 50)
    check_queries(g, o, q, 50, filter_ids=filt, threshold=8.5)
    with pytest.raises(capi.CometError):
        g.remove(int(ids[0]))                                    # already deleted
    g.flush(); o.flush()
    assert len(g) == len(o) == 4960
    check_queries(g, o, q, 50)
    x2 = rng.standard_normal((10, 40)).astype(np.float32)
    g.add(np.arange(9000, 9010), x2.copy()); o.add(np.arange(9000, 9010), x2.copy())
    check_queries(g, o, q, 50)


def test_fma_rounding_mode():
    rng = np.random.default_rng(13)
    x = rng.standard_normal((3000, 96)).astype(np.float32)
    q = rng.standard_normal((3, 96)).astype(np.float32)
    capi.check(capi.lib().cm_set_rounding(capi.ROUND_FMA)); O.set_fma(True)
    try:
        for metric in METRICS:
            g, o = build_pair(x, metric)
            check_queries(g, o, q, 20)
    finally:
        capi.check(capi.lib().cm_set_rounding(capi.ROUND_SEPARATE)); O.set_fma(False)


def test_full_size_1m_x_768_exact_path():
    """BASELINE configs[1] shape through the exact scan: 1M x 768 cosine, K=100."""
    import torch
    n, d = 1_000_000, 768
    gen = torch.Generator(device="cuda"); gen.manual_seed(20261017)
    xd = torch.randn((n, d), generator=gen, device="cuda", dtype=torch.float32)
    g = capi.FlatIndex(d, capi.COSINE)
    g.add_device(np.arange(1, n + 1, dtype=np.uint32), xd.data_ptr(), n)
    torch.cuda.synchronize()
    rng = np.random.default_rng(3)
    q = rng.standard_normal((3, d)).astype(np.float32)
    ids, sc, cnt, pos = g.search(q, k=100, path=capi.PATH_EXACT, with_pos=True)
    assert cnt.tolist() == [100, 100, 100]
    # size-independent properties: sorted, positions consistent with ids, scores reproduce from stored rows
    for i in range(3):
        assert np.all(np.diff(sc[i]) >= 0)
        assert np.array_equal(ids[i], (pos[i] + 1).astype(np.uint32))
        rows = g.get_rows(pos[i])
        qn = O.preprocess(O.COSINE, q[i])
        want = np.array([O.distance(O.COSINE, qn, r) for r in rows], np.float32)
        assert np.array_equal(bits(sc[i]), bits(want))
    # and the full oracle on one query (about 1.5 s of CPU)
    xh = xd.cpu().numpy(); del xd
    o = O.Flat(d, O.COSINE)
    o.add(np.arange(1, n + 1, dtype=np.uint32), xh)
    oi, os_ = o.search(q[0], k=100)
    assert_same_results(ids[0], sc[0], cnt[0], oi, os_, "1M x 768")


def test_merge_shards_device_matches_single_index():
    # multi-GPU exchange step on one GPU: two row shards searched separately, lists stacked as an
    # all-gather would, merged by cm_merge_shards_device; must equal the oracle on the whole corpus
    import torch
    rng = np.random.default_rng(77)
    n, d, k, nq = 5000, 48, 17, 9
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[10] = x[4000]                                   # a tie across the shard boundary
    q = rng.standard_normal((nq, d)).astype(np.float32)
    q[0] = x[10]
    ids = np.arange(1, n + 1, dtype=np.uint32)
    half = n // 2
    parts = []
    for r0, r1 in [(0, half), (half, n)]:
        g = capi.FlatIndex(d, capi.L2SQ)
        g.add(ids[r0:r1], x[r0:r1].copy())
        parts.append(g.search(q, k=k))
    g_ids = torch.from_numpy(np.stack([p[0] for p in parts]).view(np.int32)).cuda()
    g_sc = torch.from_numpy(np.stack([p[1] for p in parts])).cuda()
    g_cnt = torch.from_numpy(np.stack([p[2] for p in parts])).cuda()
    o_ids = torch.zeros((nq, k), dtype=torch.int32, device="cuda")
    o_sc = torch.zeros((nq, k), dtype=torch.float32, device="cuda")
    o_cnt = torch.zeros((nq,), dtype=torch.int64, device="cuda")
    capi.merge_shards_device(g_ids.data_ptr(), g_sc.data_ptr(), g_cnt.data_ptr(), 2, nq, k, k, o_ids.data_ptr(),
                             o_sc.data_ptr(), o_cnt.data_ptr(), out_stride=k,
                             stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    o = O.Flat(d, capi.L2SQ)
    o.add(ids, x.copy())
    for i in range(nq):
        oi, os_ = o.search(q[i], k=k)
        assert_same_results(o_ids[i].cpu().numpy().view(np.uint32), o_sc[i].cpu().numpy(), int(o_cnt[i]), oi, os_,
                            what=f"query {i}")


def test_concurrent_searches_are_reentrant():
    # flat_index_search_test.go:392-465: goroutines searching the same index concurrently (RLock holders).
    # ctypes drops the GIL during the C call, so these threads really overlap inside the library.
    import threading
    rng = np.random.default_rng(4)
    n, d, k = 30000, 64, 10
    x = rng.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    g = capi.FlatIndex(d, capi.L2SQ)
    g.add(ids, x.copy())
    o = O.Flat(d, capi.L2SQ)
    o.add(ids, x.copy())
    qs = [rng.standard_normal((m, d)).astype(np.float32) for m in (3, 70, 9, 130, 1, 64)]   # exact and tensor paths mixed
    want = [o.search_batch(q, k) for q in qs]
    errors = []

    def worker(i):
        try:
            for _ in range(5):
                gi, gs, gc = g.search(qs[i], k=k)
                oi, os_, oc = want[i]
                assert np.array_equal(gc, oc) and np.array_equal(gi, oi) and np.array_equal(bits(gs), bits(os_))
        except Exception as e:      # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(qs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_dynamic_batcher_coalesces_concurrent_single_queries():
    # many threads, one query per call (the reference's usage) -> few device batches, same answers
    import threading
    import time
    rng = np.random.default_rng(8)
    n, d, k = 70000, 64, 10
    x = rng.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    g = capi.FlatIndex(d, capi.COSINE)
    g.add(ids, x.copy())
    n_threads, per = 64, 6
    qs = rng.standard_normal((n_threads, per, d)).astype(np.float32)
    want_i, want_s, want_c = g.search(qs.reshape(-1, d), k=k, path=capi.PATH_EXACT)
    b = capi.FlatBatcher(g, max_batch=256, max_wait_us=2000)
    got = {}
    errors = []

    def worker(t):
        try:
            for j in range(per):
                got[(t, j)] = b.search(qs[t, j], k=k)
        except Exception as e:      # noqa: BLE001
            errors.append(repr(e))

    t0 = time.perf_counter()
    threads = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    dt = time.perf_counter() - t0
    assert not errors, errors
    for t in range(n_threads):
        for j in range(per):
            gi, gs = got[(t, j)]
            r = t * per + j
            assert np.array_equal(gi, want_i[r, :want_c[r]]) and np.array_equal(bits(gs), bits(want_s[r, :want_c[r]]))
    batches, requests = b.stats()
    assert requests == n_threads * per and batches < requests / 4, (batches, requests)
    with pytest.raises(capi.CometError) as e:
        b.search(np.zeros(d, np.float32), k=k)          # zero query under cosine: the batch reports ErrZeroVector
    assert e.value.code == capi.ERR_ZERO_VECTOR
    b.close()
    print(f"batcher: {requests} single-query calls in {batches} device batches, {requests / dt:.0f} q/s")
