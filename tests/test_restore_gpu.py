"""Restore path: the byte formats stay host code, the device side must take back the decoded
state WITHOUT re-deriving it -- stored rows, list membership and codes are imported as they are and the
restored index must answer exactly like the one that was built."""
import numpy as np
import pytest

from comet_b200 import capi
from tests.parity import bits

pytestmark = pytest.mark.gpu


def same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(bits(a[1]), bits(b[1])) and np.array_equal(a[2], b[2])


def test_flat_restore_cosine_rows_are_not_renormalised():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3000, 48)).astype(np.float32)
    ids = np.arange(1, 3001, dtype=np.uint32)
    a = capi.FlatIndex(48, capi.COSINE)
    a.add(ids, x.copy())
    stored = a.get_rows(np.arange(3000))
    b = capi.FlatIndex(48, capi.COSINE)
    b.load_rows(ids, stored)
    assert np.array_equal(bits(b.get_rows(np.arange(3000))), bits(stored))
    q = rng.standard_normal((9, 48)).astype(np.float32)
    assert same(a.search(q, k=10), b.search(q, k=10))


def test_ivf_pq_ivfpq_restore():
    rng = np.random.default_rng(2)
    n, d = 4000, 32
    x = rng.standard_normal((n, d)).astype(np.float32) + 0.3
    ids = np.arange(1, n + 1, dtype=np.uint32)
    q = rng.standard_normal((7, d)).astype(np.float32)
    # IVF
    a = capi.IVFIndex(d, 16, capi.COSINE)
    a.train(x[:1000].copy())
    lists = a.add(ids, x.copy())
    b = capi.IVFIndex(d, 16, capi.COSINE)
    b.set_centroids(a.centroids())
    order = np.argsort(lists, kind="stable")                    # the IVFX format stores list by list
    b.load_lists(ids[order], a.get_rows(order), lists[order])
    assert same(a.search(q, k=10, nprobes=4), b.search(q, k=10, nprobes=4))
    # PQ
    a = capi.PQIndex(d, capi.L2, 8, 4)
    a.train(x[:1000].copy())
    a.add(ids, x.copy())
    b = capi.PQIndex(d, capi.L2, 8, 4)
    b.set_codebooks(a.codebooks())
    b.load_codes(ids, a.codes())
    assert same(a.search(q, k=10), b.search(q, k=10))
    # IVFPQ
    a = capi.IVFPQIndex(d, capi.L2, 8, 8, 4)
    a.train(x[:1000].copy())
    lists = a.add(ids, x.copy())
    b = capi.IVFPQIndex(d, capi.L2, 8, 8, 4)
    b.set_trained(*a.trained_state())
    order = np.argsort(lists, kind="stable")
    b.load_codes(ids[order], a.codes()[order], lists[order])
    assert same(a.search(q, k=10, nprobes=3), b.search(q, k=10, nprobes=3))
