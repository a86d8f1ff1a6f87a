"""GPU parity of the build side: Train on the device must reproduce the reference's
deterministic k-means (clustering.go:119-239) bit for bit -- centroids and codebooks are compared with
the CPU oracle's, float bits and all."""
import numpy as np
import pytest

from comet_b200 import capi
from oracle import oracle_py as O
from tests.parity import bits

pytestmark = pytest.mark.gpu


def data(n, d, seed, shift=0.0):
    rng = np.random.default_rng(seed)
    # a few separated blobs so that k-means runs several iterations and clusters differ in size
    centers = rng.standard_normal((7, d)).astype(np.float32) * 3
    x = centers[rng.integers(0, 7, n)] + rng.standard_normal((n, d)).astype(np.float32) + np.float32(shift)
    return x.astype(np.float32), rng


@pytest.mark.parametrize("metric", [capi.L2, capi.L2SQ, capi.COSINE])
def test_ivf_train_bit_exact(metric):
    x, rng = data(3000, 40, 5 + metric, 0.5)
    o = O.IVF(40, 23, metric)
    o.train(x.copy())
    g = capi.IVFIndex(40, 23, metric)
    g.train(x.copy())
    assert np.array_equal(bits(g.centroids()), bits(o.centroids()))
    ids = np.arange(1, 3001, dtype=np.uint32)
    o.add(ids, x.copy())
    g.add(ids, x.copy())
    q = rng.standard_normal((5, 40)).astype(np.float32)
    gi, gs, gc = g.search(q, k=10, nprobes=4)
    for i in range(5):
        oi, os_ = o.search(q[i], k=10, nprobes=4)
        assert np.array_equal(gi[i, :gc[i]], oi) and np.array_equal(bits(gs[i, :gc[i]]), bits(os_))
    with pytest.raises(capi.CometError) as e:
        capi.IVFIndex(40, 23, metric).train(x[:10].copy())
    assert e.value.code == capi.ERR_TOO_FEW and "need at least 23 training vectors for 23 clusters (got 10)" in e.value.msg


def test_pq_train_bit_exact():
    x, rng = data(2500, 32, 11)
    o = O.PQ(32, capi.L2, 8, 5)
    o.train(x.copy())
    g = capi.PQIndex(32, capi.L2, 8, 5)
    g.train(x.copy())
    assert np.array_equal(bits(g.codebooks()), bits(o.codebooks()))
    with pytest.raises(capi.CometError) as e:
        capi.PQIndex(32, capi.L2, 8, 5).train(x[:20].copy())
    assert e.value.code == capi.ERR_TOO_FEW and "need at least 32 vectors for training" in e.value.msg


@pytest.mark.parametrize("metric", [capi.L2, capi.COSINE])
def test_ivfpq_train_bit_exact(metric):
    x, rng = data(2000, 24, 21 + metric, 0.3)
    o = O.IVFPQ(24, metric, 12, 6, 4)
    o.train(x.copy())
    g = capi.IVFPQIndex(24, metric, 12, 6, 4)
    g.train(x.copy())
    c, cb = g.trained_state()
    assert np.array_equal(bits(c), bits(o.centroids()))
    assert np.array_equal(bits(cb), bits(o.codebooks()))
    with pytest.raises(capi.CometError) as e:
        capi.IVFPQIndex(24, metric, 12, 6, 4).train(x[:100].copy())
    assert e.value.code == capi.ERR_TOO_FEW and "need at least 120 vectors for training" in e.value.msg
