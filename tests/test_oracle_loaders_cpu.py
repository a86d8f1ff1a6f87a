"""The oracle's ReadFrom-style loaders (used by the config-size GPU parity tests to put device-built state into
the checker) restore exactly the state an oracle index built the normal way holds."""
import numpy as np

from oracle import oracle_py as O


def test_ivfpq_load_codes_equals_add():
    rng = np.random.default_rng(3)
    n, d = 3000, 32
    x = rng.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    a = O.IVFPQ(d, O.L2, 16, 8, 4)
    a.train(x[:1000].copy())
    a.add(ids, x.copy())
    # stored state, list by list
    l_ids, l_codes, l_of = [], [], []
    for l, (i_, c_) in enumerate(a.lists()):
        l_ids.append(i_); l_codes.append(c_); l_of.append(np.full(len(i_), l, np.int32))
    b = O.IVFPQ(d, O.L2, 16, 8, 4)
    b.set_trained(a.centroids(), a.codebooks())
    b.load_codes(np.concatenate(l_ids), np.concatenate(l_codes), np.concatenate(l_of))
    q = rng.standard_normal((5, d)).astype(np.float32)
    for i in range(5):
        ai, as_ = a.search(q[i], k=20, nprobes=5)
        bi, bs = b.search(q[i], k=20, nprobes=5)
        assert np.array_equal(ai, bi) and np.array_equal(as_.view(np.uint32), bs.view(np.uint32))


def test_hnsw_load_graph_equals_built_graph():
    rng = np.random.default_rng(4)
    n, d = 400, 16
    x = rng.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    levels = O.hnsw_random_levels(n, 8, 7)
    a = O.HNSW(d, O.COSINE, 8, 40, 30)
    a.add(ids, x.copy(), levels)
    e_ids, e_levels, e_rows, layers = a.export()
    off, chunks = [0], []
    for s in range(n):
        for layer in range(int(e_levels[s]) + 1):
            lo, nb = layers[layer]
            seg = nb[lo[s]:lo[s + 1]]
            chunks.append(seg); off.append(off[-1] + len(seg))
    b = O.HNSW(d, O.COSINE, 8, 40, 30)
    b.load_graph(e_ids, e_rows, e_levels, np.asarray(off, np.int64), np.concatenate(chunks), a.entry_point, a.max_level)
    q = rng.standard_normal((6, d)).astype(np.float32)
    for i in range(6):
        ai, as_ = a.search(q[i], k=10)
        wa = O.HNSW.last_counters()
        bi, bs = b.search(q[i], k=10)
        wb = O.HNSW.last_counters()
        assert np.array_equal(ai, bi) and np.array_equal(as_.view(np.uint32), bs.view(np.uint32)) and wa == wb
