"""Row-sharded FlatIndex behind the C ABI (cm_flat_sharded_*): one process drives every shard.

The box the `-m gpu` tier runs on has one GPU, so the shards of these tests share a device (the library takes
the same device more than once); with more GPUs visible the same tests spread the shards over them.  Every result
is compared bit-exactly with the oracle's SINGLE FlatIndex over the whole corpus: the merge must reproduce the
reference's (score, scan position) order (flat_index_search.go:277-291), also for ties that straddle shards.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from comet_b200 import capi  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from tests.parity import assert_same_results, bits  # noqa: E402


def _devices(n_shards):
    import torch
    n_dev = torch.cuda.device_count()
    return [r % n_dev for r in range(n_shards)]


def _check(g, o, q, k, **kw):
    ids, sc, cnt = g.search(q, k=k, **kw)
    for i in range(len(q)):
        oi, os_ = o.search(q[i], k=k, **{a: b for a, b in kw.items() if a != "path"})
        assert_same_results(ids[i], sc[i], cnt[i], oi, os_, what=f"query {i}")
    return ids, sc, cnt


@pytest.mark.parametrize("metric", [capi.L2SQ, capi.L2, capi.COSINE])
def test_sharded_small_exact_path_with_ties_across_shards(metric):
    rng = np.random.default_rng(5)
    n, d, k, nq, per = 5000, 48, 17, 9, 1300            # 4 shards, the last one ragged (1100 rows)
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[10] = x[4000]                                      # ties across shards 0 / 3 and 1 / 2
    x[1400] = x[2700]
    q = rng.standard_normal((nq, d)).astype(np.float32)
    q[0], q[1] = x[10], x[1400]
    ids = np.arange(1, n + 1, dtype=np.uint32)
    g = capi.ShardedFlatIndex(d, metric, _devices(4), per)
    g.add(ids, x.copy())
    assert len(g) == n and [g.shard_size(r) for r in range(4)] == [1300, 1300, 1300, 1100]
    o = O.Flat(d, metric)
    o.add(ids, x.copy())
    _check(g, o, q, k)
    _check(g, o, q, 1)
    _check(g, o, q, 1500)                                # k larger than a shard
    _check(g, o, q[:1], 0)                               # k <= 0: everything, in order
    _check(g, o, q, 5)                                   # a larger batch with a smaller k after it: per-query buffers regrow


def test_sharded_duplicate_rows_everywhere():
    # every row equals one of 7 prototypes: all scores tie in big groups, order must be scan order across shards
    rng = np.random.default_rng(6)
    d, per = 32, 200
    protos = rng.standard_normal((7, d)).astype(np.float32)
    x = protos[rng.integers(0, 7, size=1000)]
    ids = np.arange(1, 1001, dtype=np.uint32)
    g = capi.ShardedFlatIndex(d, capi.L2SQ, _devices(5), per)
    g.add(ids, x.copy())
    o = O.Flat(d, capi.L2SQ)
    o.add(ids, x.copy())
    _check(g, o, protos[:3].copy(), 333)


@pytest.mark.parametrize("metric", [capi.L2, capi.COSINE])
def test_sharded_tensor_path(metric):
    # shards large enough for the tcgen05 candidate pass (>= 65,536 rows, >= 64 queries)
    rng = np.random.default_rng(7)
    per, d, k, nq = 70000, 128, 50, 96
    n = 2 * per + 30000                                  # third shard is below the tensor-path size: exact scan there
    x = rng.standard_normal((n, d)).astype(np.float32)
    x[5] = x[per + 9]
    x[per - 1] = x[2 * per + 17]
    q = rng.standard_normal((nq, d)).astype(np.float32)
    q[0], q[1] = x[5], x[per - 1]
    ids = np.arange(1, n + 1, dtype=np.uint32)
    g = capi.ShardedFlatIndex(d, metric, _devices(3), per)
    g.add(ids, x.copy())
    o = O.Flat(d, metric)
    o.add(ids, x.copy())
    O.set_threads(16)
    o_ids, o_sc, o_cnt = o.search_batch(q, k)
    ids_g, sc_g, cnt_g = g.search(q, k=k)
    assert np.array_equal(cnt_g, o_cnt)
    assert np.array_equal(ids_g, o_ids[:, :k])
    assert np.array_equal(bits(sc_g), bits(o_sc[:, :k]))
    st = g.last_stats()
    assert st["path_used"] == capi.PATH_TENSOR and st["fallback_queries"] == 0
    # forcing the exact path everywhere gives the same bits
    ids_e, sc_e, cnt_e = g.search(q, k=k, path=capi.PATH_EXACT)
    assert np.array_equal(ids_e, ids_g) and np.array_equal(bits(sc_e), bits(sc_g))


def test_sharded_remove_flush_threshold_filter_and_more_adds():
    rng = np.random.default_rng(8)
    n, d, per = 3000, 40, 1000
    x = rng.standard_normal((n + 500, d)).astype(np.float32)
    q = rng.standard_normal((5, d)).astype(np.float32)
    ids = np.arange(1, n + 501, dtype=np.uint32)
    g = capi.ShardedFlatIndex(d, capi.L2SQ, _devices(4), per)
    o = O.Flat(d, capi.L2SQ)
    g.add(ids[:n], x[:n].copy())
    o.add(ids[:n], x[:n].copy())
    for dead in (7, 999, 1000, 1001, 2500):
        g.remove(dead)
        o.remove(dead)
    with pytest.raises(capi.CometError) as e:
        g.remove(999)                                    # already deleted
    assert e.value.code == capi.ERR_NOT_FOUND
    _check(g, o, q, 25)
    thr = float(np.sort(((x[:n] - q[0]) ** 2).sum(1))[40])
    _check(g, o, q, 100, threshold=thr)
    filt = np.array([5, 7, 1500, 2999, 3000, 77777], dtype=np.uint32)
    _check(g, o, q, 10, filter_ids=filt)
    g.flush()
    o.flush()
    assert len(g) == n - 5 and [g.shard_size(r) for r in range(4)] == [997, 999, 999, 0]
    # rows added after a flush go behind everything (scan order), not into the gaps of earlier shards
    g.add(ids[n:], x[n:].copy())
    o.add(ids[n:], x[n:].copy())
    assert [g.shard_size(r) for r in range(4)] == [997, 999, 1000, 499]    # appended behind the last row
    _check(g, o, q, 60)


def test_sharded_errors_and_empty_index():
    d = 16
    g = capi.ShardedFlatIndex(d, capi.COSINE, _devices(2), 10)
    q = np.ones((2, d), np.float32)
    ids, sc, cnt = g.search(q, k=5)
    assert cnt.tolist() == [0, 0]
    qz = q.copy()
    qz[1] = 0
    with pytest.raises(capi.CometError) as e:            # Preprocess fails before the index is looked at
        g.search(qz, k=5)
    assert e.value.code == capi.ERR_ZERO_VECTOR
    x = np.random.default_rng(1).standard_normal((20, d)).astype(np.float32)
    g.add(np.arange(1, 21, dtype=np.uint32), x.copy())
    with pytest.raises(capi.CometError) as e:
        g.add(np.array([99], np.uint32), x[:1].copy())
    assert e.value.code == capi.ERR_UNSUPPORTED          # both shards are full
    with pytest.raises(capi.CometError) as e:
        g.search(qz, k=5)
    assert e.value.code == capi.ERR_ZERO_VECTOR
    with pytest.raises(capi.CometError) as e:
        g.search(np.ones((1, d + 1), np.float32), k=5)
    assert e.value.code == capi.ERR_DIM_MISMATCH
    # a zero ROW under cosine stops the batch at that row, like n successive Adds
    g2 = capi.ShardedFlatIndex(d, capi.COSINE, _devices(2), 10)
    xb = x[:15].copy()
    xb[12] = 0
    with pytest.raises(capi.CometError) as e:
        g2.add(np.arange(1, 16, dtype=np.uint32), xb)
    assert e.value.code == capi.ERR_ZERO_VECTOR and len(g2) == 12


def test_sharded_device_entry_point_reports_exchange_bytes():
    import torch
    rng = np.random.default_rng(9)
    n, d, k, nq, per = 4000, 64, 20, 33, 1000
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    devs = _devices(4)
    g = capi.ShardedFlatIndex(d, capi.L2, devs, per)
    g.add(ids, x.copy())
    torch.cuda.set_device(devs[0])
    qd = torch.from_numpy(q).cuda()
    o_ids = torch.zeros((nq, k + 3), dtype=torch.int32, device="cuda")
    o_sc = torch.zeros((nq, k + 3), dtype=torch.float32, device="cuda")
    o_cnt = torch.zeros((nq,), dtype=torch.int64, device="cuda")
    g.search_device(qd.data_ptr(), nq, k, o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(), k + 3,
                    stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    o = O.Flat(d, capi.L2)
    o.add(ids, x.copy())
    for i in range(nq):
        oi, os_ = o.search(q[i], k=k)
        assert_same_results(o_ids[i, :k].cpu().numpy().view(np.uint32), o_sc[i, :k].cpu().numpy(), int(o_cnt[i]), oi, os_)
    remote = sum(1 for r in devs if r != devs[0])
    assert g.exchange_bytes() == remote * (nq * d * 4 + nq * k * 8 + nq * 8)


# ---- PQIndex row shards (cm_pq_sharded_*) ---------------------------------------------------------------------------
@pytest.mark.parametrize("metric", [capi.L2, capi.COSINE])
def test_pq_row_shards_match_the_single_index(metric):
    rng = np.random.default_rng(40 + metric)
    n, d, M, nbits, per = 5000, 32, 8, 4, 1300            # 16 codewords per sub-quantiser: ADC scores tie constantly
    x = rng.standard_normal((n, d)).astype(np.float32) + (0.2 if metric == capi.COSINE else 0.0)
    x[10] = x[4000]
    q = rng.standard_normal((7, d)).astype(np.float32)
    q[0] = x[10]
    ids = np.arange(1, n + 1, dtype=np.uint32)
    o = O.PQ(d, metric, M, nbits)
    o.train(x[:800].copy())
    g = capi.ShardedPQIndex(d, metric, M, nbits, _devices(4), per)
    with pytest.raises(capi.CometError) as e:
        g.add(ids[:5], x[:5].copy())
    assert e.value.code == capi.ERR_NOT_TRAINED
    g.set_codebooks(o.codebooks())
    o.add(ids, x.copy())
    g.add(ids, x.copy())
    assert len(g) == n
    _check(g, o, q, 10)
    _check(g, o, q, 300)
    _check(g, o, q[:2], 0)
    _check(g, o, q[:3], 40, filter_ids=np.arange(2, n, 3, dtype=np.uint32))
    for dead in (11, 1300, 1301, 4999):
        g.remove(dead)
        o.remove(dead)
    _check(g, o, q, 25)
    g.flush()
    o.flush()
    assert len(g) == n - 4
    _check(g, o, q, 25)
    # trained on the device: codebooks equal on every shard, answers equal to a single device-trained index
    g2 = capi.ShardedPQIndex(d, metric, M, nbits, _devices(3), 2000)
    g2.train(x[:800].copy())
    o2 = O.PQ(d, metric, M, nbits)
    o2.set_codebooks(g2.codebooks())
    g2.add(ids, x.copy())
    o2.add(ids, x.copy())
    _check(g2, o2, q, 15)
    if metric == capi.COSINE:
        qz = q[:2].copy()
        qz[1] = 0
        with pytest.raises(capi.CometError) as e:
            g2.search(qz, k=3)
        assert e.value.code == capi.ERR_ZERO_VECTOR
