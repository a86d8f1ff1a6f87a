"""WriteTo / ReadFrom through the C ABI (cm_*_save / cm_*_load / *_file) and HNSWIndex.Flush (cm_hnsw_flush).

For every index type: the product's bytes must EQUAL the oracle writer's bytes for the same state (the oracle
writer is pinned field by field in tests/test_wire_cpu.py), a fresh index loaded from them must answer bit-identically
to the index that was saved and to the oracle, gzip segment files (`vector_%06d.bin.gz`, storage_provider.go:163-166)
must round-trip both ways, and a bad stream must leave the index untouched."""
import gzip

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from comet_b200 import capi  # noqa: E402
from oracle import oracle_py as O  # noqa: E402
from oracle import wire_py as W  # noqa: E402
from tests.parity import assert_same_results, bits  # noqa: E402


def same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1])) and np.array_equal(a[2], b[2])


def vs_oracle(g, o, q, **kw):
    ids, sc, cnt = g.search(q, **kw)
    for i in range(len(q)):
        oi, os_ = o.search(q[i], **kw)
        assert_same_results(ids[i], sc[i], cnt[i], oi, os_, what=f"query {i}")


def data(seed, n, d):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d)).astype(np.float32) + 0.25
    q = rng.standard_normal((6, d)).astype(np.float32)
    return x, np.arange(1, n + 1, dtype=np.uint32), q


@pytest.mark.parametrize("metric", [capi.COSINE, capi.L2])
def test_flat_save_load(metric, tmp_path):
    x, ids, q = data(1, 2500, 40)
    a, o = capi.FlatIndex(40, metric), O.Flat(40, metric)
    a.add(ids, x.copy())
    o.add(ids, x.copy())
    for dead in (3, 1200, 2500):
        a.remove(dead)
        o.remove(dead)
    blob = capi.save_bytes("flat", a.h)
    assert len(a) == 2497                                        # WriteTo flushes first (flat_index.go:367-370)
    o.flush()
    assert blob == W.write_flat(40, metric, o.ids(), o.rows())   # byte for byte the reference's stream
    b = capi.FlatIndex(40, metric)
    b.add(ids[:10], x[:10].copy())                               # ReadFrom REPLACES whatever the index held
    assert capi.load_bytes("flat", b.h, blob) == len(blob)
    assert len(b) == 2497 and same(a.search(q, k=20), b.search(q, k=20))
    vs_oracle(b, o, q, k=20)
    assert capi.save_bytes("flat", b.h) == blob
    # segment files: gzip and plain, both directions
    gz = tmp_path / "vector_000001.bin.gz"
    capi.save_file("flat", a.h, gz)
    assert gzip.open(gz, "rb").read() == blob
    plain = tmp_path / "vector_000002.bin"
    capi.save_file("flat", a.h, plain)
    assert plain.read_bytes() == blob
    theirs = tmp_path / "theirs.bin.gz"
    with gzip.open(theirs, "wb") as f:
        f.write(blob)
    for path in (gz, plain, theirs):
        c = capi.FlatIndex(40, metric)
        capi.load_file("flat", c.h, path)
        assert same(a.search(q, k=20), c.search(q, k=20))
    # a stream whose deleted set is not empty (the reference never writes one, its reader accepts it)
    marked = W.write_flat(40, metric, o.ids(), o.rows(), deleted_ids=[7, 8, 70000])
    c = capi.FlatIndex(40, metric)
    capi.load_bytes("flat", c.h, marked)
    o.remove(7)
    o.remove(8)
    vs_oracle(c, o, q, k=20)


def test_flat_bad_streams_leave_the_index_untouched():
    x, ids, q = data(2, 300, 16)
    a = capi.FlatIndex(16, capi.L2SQ)
    a.add(ids, x.copy())
    before = a.search(q, k=5)
    good = capi.save_bytes("flat", a.h)
    cases = [
        (b"FLAX" + good[4:], capi.ERR_INVALID_ARG, "invalid magic number: expected 'FLAT', got 'FLAX'"),
        (good[:4] + (2).to_bytes(4, "little") + good[8:], capi.ERR_UNSUPPORTED, "unsupported version: 2"),
        (good[:len(good) // 2], capi.ERR_INVALID_ARG, "unexpected EOF"),
        (W.write_flat(17, capi.L2SQ, ids[:1], np.ones((1, 17), np.float32)), capi.ERR_DIM_MISMATCH,
         "dimension mismatch: index has dim=16, serialized data has dim=17"),
        (W.write_flat(16, capi.COSINE, ids[:1], np.ones((1, 16), np.float32)), capi.ERR_INVALID_ARG,
         "distance kind mismatch: index uses 'l2_squared', serialized data uses 'cosine'"),
    ]
    for blob, code, text in cases:
        with pytest.raises(capi.CometError) as e:
            capi.load_bytes("flat", a.h, blob)
        assert e.value.code == code and text in e.value.msg
        assert same(before, a.search(q, k=5))
    # size query, then a buffer that is too small
    import ctypes as C
    n = C.c_int64(0)
    capi.check(capi.lib().cm_flat_save(a.h, None, 0, C.byref(n)))
    assert n.value == len(good)
    small = np.zeros(10, np.uint8)
    assert capi.lib().cm_flat_save(a.h, capi.ptr(small, capi.u8p), 10, C.byref(n)) == capi.ERR_BUFFER_TOO_SMALL and n.value == len(good)
    # empty index
    e = capi.FlatIndex(16, capi.L2SQ)
    blob = capi.save_bytes("flat", e.h)
    assert blob == W.write_flat(16, capi.L2SQ, np.zeros(0, np.uint32), np.zeros((0, 16), np.float32))
    capi.load_bytes("flat", a.h, blob)
    assert len(a) == 0


def test_ivf_save_load(tmp_path):
    x, ids, q = data(3, 3000, 32)
    a, o = capi.IVFIndex(32, 12, capi.COSINE), O.IVF(32, 12, capi.COSINE)
    untrained = capi.save_bytes("ivf", a.h)
    assert untrained == W.write_ivf(32, capi.COSINE, 12, None, [(np.zeros(0, np.uint32), np.zeros((0, 32), np.float32))] * 12)
    a.train(x[:600].copy())
    o.set_centroids(a.centroids())
    a.add(ids, x.copy())
    o.add(ids, x.copy())
    for dead in (5, 77, 2999):
        a.remove(dead)
        o.remove(dead)
    blob = capi.save_bytes("ivf", a.h)
    o.flush()
    assert blob == W.write_ivf(32, capi.COSINE, 12, o.centroids(), o.lists())
    b = capi.IVFIndex(32, 12, capi.COSINE)
    assert capi.load_bytes("ivf", b.h, blob) == len(blob)
    assert len(b) == 2997
    for np_ in (1, 4, 12):
        assert same(a.search(q, k=15, nprobes=np_), b.search(q, k=15, nprobes=np_))
        vs_oracle(b, o, q, k=15, nprobes=np_)
    assert capi.save_bytes("ivf", b.h) == blob
    capi.save_file("ivf", b.h, tmp_path / "vector_000003.bin.gz")
    c = capi.IVFIndex(32, 12, capi.COSINE)
    capi.load_file("ivf", c.h, tmp_path / "vector_000003.bin.gz")
    assert same(a.search(q, k=15, nprobes=4), c.search(q, k=15, nprobes=4))
    # a loaded index keeps working: more adds, then search
    more = np.arange(5001, 5101, dtype=np.uint32)
    c.add(more, x[:100].copy() * 1.5)
    o.add(more, x[:100].copy() * 1.5)
    vs_oracle(c, o, q, k=15, nprobes=4)
    with pytest.raises(capi.CometError) as e:
        other = capi.IVFIndex(32, 13, capi.COSINE)
        capi.load_bytes("ivf", other.h, blob)
    assert "nlist mismatch: index has nlist=13, serialized data has nlist=12" in e.value.msg
    capi.load_bytes("ivf", c.h, untrained)                       # back to an untrained, empty index
    assert len(c) == 0
    with pytest.raises(capi.CometError) as e:
        c.search(q, k=3)
    assert e.value.code == capi.ERR_NOT_TRAINED


def test_pq_and_ivfpq_save_load(tmp_path):
    x, ids, q = data(4, 3000, 32)
    # PQ
    a, o = capi.PQIndex(32, capi.L2, 8, 4), O.PQ(32, capi.L2, 8, 4)
    a.train(x[:800].copy())
    o.set_codebooks(a.codebooks())
    a.add(ids, x.copy())
    o.add(ids, x.copy())
    for dead in (9, 1500):
        a.remove(dead)
        o.remove(dead)
    blob = capi.save_bytes("pq", a.h)
    o.flush()
    assert blob == W.write_pq(32, capi.L2, 8, 4, o.codebooks(), o.ids(), o.codes())
    b = capi.PQIndex(32, capi.L2, 8, 4)
    assert capi.load_bytes("pq", b.h, blob) == len(blob) and len(b) == 2998
    assert same(a.search(q, k=25), b.search(q, k=25))
    vs_oracle(b, o, q, k=25)
    with pytest.raises(capi.CometError) as e:
        other = capi.PQIndex(32, capi.L2, 4, 4)
        capi.load_bytes("pq", other.h, blob)
    assert "parameter M mismatch: index has M=4, serialized data has M=8" in e.value.msg
    capi.save_file("pq", b.h, tmp_path / "pq.bin.gz")
    assert gzip.open(tmp_path / "pq.bin.gz", "rb").read() == blob
    # IVFPQ
    a, o = capi.IVFPQIndex(32, capi.L2, 10, 8, 4), O.IVFPQ(32, capi.L2, 10, 8, 4)
    a.train(x[:1000].copy())
    o.set_trained(*a.trained_state())
    a.add(ids, x.copy())
    o.add(ids, x.copy())
    for dead in (1, 2, 2222):
        a.remove(dead)
        o.remove(dead)
    blob = capi.save_bytes("ivfpq", a.h)
    o.flush()
    assert blob == W.write_ivfpq(32, capi.L2, 10, 8, 4, o.centroids(), o.codebooks(), o.lists())
    b = capi.IVFPQIndex(32, capi.L2, 10, 8, 4)
    assert capi.load_bytes("ivfpq", b.h, blob) == len(blob) and len(b) == 2997
    for np_ in (1, 3, 10):
        assert same(a.search(q, k=25, nprobes=np_), b.search(q, k=25, nprobes=np_))
        vs_oracle(b, o, q, k=25, nprobes=np_)
    assert capi.save_bytes("ivfpq", b.h) == blob
    capi.save_file("ivfpq", b.h, tmp_path / "ivfpq.bin")
    c = capi.IVFPQIndex(32, capi.L2, 10, 8, 4)
    capi.load_file("ivfpq", c.h, tmp_path / "ivfpq.bin")
    assert same(a.search(q, k=25, nprobes=3), c.search(q, k=25, nprobes=3))
    with pytest.raises(capi.CometError) as e:
        capi.load_bytes("ivfpq", c.h, blob[:-40])
    assert "unexpected EOF" in e.value.msg
    assert same(a.search(q, k=25, nprobes=3), c.search(q, k=25, nprobes=3))     # untouched


def _oracle_nodes(o):
    ids, levels, rows, layers = o.export()
    nodes = []
    for s in range(len(ids)):
        edges = [layers[l][1][layers[l][0][s]:layers[l][0][s + 1]] for l in range(int(levels[s]) + 1)]
        nodes.append((int(ids[s]), int(levels[s]), rows[s], edges))
    return nodes


def _load_oracle(nodes, entry, max_level, d, metric, m, efc, efs):
    o = O.HNSW(d, metric, m, efc, efs)
    ids = np.array([n[0] for n in nodes], np.uint32)
    rows = np.stack([n[2] for n in nodes]) if nodes else np.zeros((0, d), np.float32)
    levels = np.array([n[1] for n in nodes], np.int32)
    offs, chunks = [0], []
    for n in nodes:
        for e in n[3]:
            chunks.append(np.asarray(e, np.uint32))
            offs.append(offs[-1] + len(e))
    edge_ids = np.concatenate(chunks) if chunks else np.zeros(0, np.uint32)
    o.load_graph(ids, rows, levels, np.asarray(offs, np.int64), edge_ids, entry, max_level)
    return o


@pytest.mark.parametrize("metric", [capi.L2, capi.COSINE])
def test_hnsw_flush_and_save_load(metric, tmp_path):
    rng = np.random.default_rng(11)
    n, d, m, efc, efs = 400, 24, 6, 40, 30
    x = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal((8, d)).astype(np.float32)
    ids = np.arange(10, 10 + n, dtype=np.uint32)
    levels = O.hnsw_random_levels(n, m, 5)
    levels[0] = max(int(levels.max()), 2)            # the first node is the entry point (never promoted, hnsw_index.go:266-284)
    levels[7] = levels[0]                            # another node at maxLevel: strategy 1 of Flush phase 2
    a, o = capi.HNSWIndex(d, metric, m, efc, efs), O.HNSW(d, metric, m, efc, efs)
    a.add(ids, x.copy(), levels)
    o.add(ids, x.copy(), levels)
    entry0 = o.entry_point
    assert entry0 == 10
    dead = [int(entry0), 11, 12, 200, 10 + n - 1]    # the entry point and a few of its neighbourhood
    for v in dead:
        a.remove(v)
        o.remove(v)
    vs_oracle(a, o, q, k=5)                          # soft-deleted state still matches
    want_nodes, want_entry, want_ml = W.hnsw_flush(_oracle_nodes(o), entry0, o.max_level, dead)
    assert want_entry == 17 and want_ml == int(levels[0])
    a.flush()
    assert len(a) == n - len(dead) and a.max_level() == want_ml
    lv, eo, ei, entry, ml = a.export_graph()
    assert entry == want_entry and ml == want_ml and lv.tolist() == [w[1] for w in want_nodes]
    flat_edges = np.concatenate([np.asarray(e, np.uint32) for w in want_nodes for e in w[3]])
    assert np.array_equal(ei, flat_edges)
    o2 = _load_oracle(want_nodes, want_entry, want_ml, d, metric, m, efc, efs)
    vs_oracle(a, o2, q, k=5)
    with pytest.raises(capi.CometError):
        a.remove(11)                                 # gone for good
    # a second flush with nothing deleted is a no-op; deleting everything empties the index
    a.flush()
    assert len(a) == n - len(dead)
    # WriteTo / ReadFrom
    blob = capi.save_bytes("hnsw", a.h)
    assert blob == W.write_hnsw(d, metric, m, efc, efs, want_ml, want_entry, want_nodes)
    b = capi.HNSWIndex(d, metric, m, efc, efs)
    assert capi.load_bytes("hnsw", b.h, blob) == len(blob)
    assert same(a.search(q, k=5), b.search(q, k=5))
    vs_oracle(b, o2, q, k=5)
    assert capi.save_bytes("hnsw", b.h) == blob
    # node order in the stream is free (the reference writes a Go map): a shuffled stream loads to the same answers
    perm = rng.permutation(len(want_nodes))
    shuffled = W.write_hnsw(d, metric, m, efc, efs, want_ml, want_entry, [want_nodes[i] for i in perm])
    c = capi.HNSWIndex(d, metric, m, efc, efs)
    capi.load_bytes("hnsw", c.h, shuffled)
    assert same(a.search(q, k=5), c.search(q, k=5))
    capi.save_file("hnsw", a.h, tmp_path / "vector_000009.bin.gz")
    assert gzip.open(tmp_path / "vector_000009.bin.gz", "rb").read() == blob
    e = capi.HNSWIndex(d, metric, m, efc, efs)
    capi.load_file("hnsw", e.h, tmp_path / "vector_000009.bin.gz")
    assert same(a.search(q, k=5), e.search(q, k=5))
    with pytest.raises(capi.CometError) as err:
        other = capi.HNSWIndex(d, metric, m + 1, efc, efs)
        capi.load_bytes("hnsw", other.h, blob)
    assert f"m parameter mismatch: index has m={m + 1}, serialized data has m={m}" in err.value.msg
    # everything deleted: Flush leaves an empty index (entry 0, maxLevel -1) that serialises and searches
    for w in want_nodes:
        e.remove(w[0])
    e.flush()
    assert len(e) == 0 and e.max_level() == -1
    assert capi.save_bytes("hnsw", e.h) == W.write_hnsw(d, metric, m, efc, efs, -1, 0, [])
    assert e.search(q, k=5)[2].tolist() == [0] * len(q)
