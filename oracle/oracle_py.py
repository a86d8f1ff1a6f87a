"""ctypes binding of the CPU parity oracle (oracle/comet_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under comet_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "libcomet_oracle.so")

L2, L2SQ, COSINE = 0, 1, 2
METRICS = {"l2": L2, "l2_squared": L2SQ, "cosine": COSINE}
AGG = {"sum": 0, "max": 1, "mean": 2}

ERR_ZERO_VECTOR, ERR_DIM, ERR_NOT_TRAINED, ERR_TOO_FEW, ERR_ARG, ERR_NOT_FOUND, ERR_UNSUPPORTED = (
    -1, -2, -3, -4, -5, -6, -7)


class OracleError(RuntimeError):
    def __init__(self, code, what=""):
        super().__init__(f"oracle error {code} {what}")
        self.code = code


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "comet_oracle.c")
    hdr = os.path.join(_HERE, "comet_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int)
i64p = C.POINTER(C.c_long)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp = C.c_void_p
    sig = {
        "co_set_fma": (None, [C.c_int]),
        "co_get_fma": (C.c_int, []),
        "co_set_threads": (None, [C.c_int]),
        "co_distance": (C.c_float, [C.c_int, f32p, f32p, C.c_int]),
        "co_normalize": (C.c_int, [f32p, f32p, C.c_int]),
        "co_preprocess": (C.c_int, [C.c_int, f32p, f32p, C.c_int]),
        "co_norm": (C.c_float, [f32p, C.c_int]),
        "co_sanitize_k": (C.c_long, [C.c_long, C.c_long]),
        "co_aggregate": (C.c_long, [C.c_int, u32p, f32p, C.c_long, u32p, f32p]),
        "co_autocut": (C.c_long, [f32p, C.c_long, C.c_int]),
        "co_kmeans": (C.c_int, [f32p, C.c_long, C.c_int, C.c_long, C.c_int, C.c_int, C.c_int, f32p, i32p]),
        "co_nearest_centroid": (C.c_int, [f32p, f32p, C.c_int, C.c_int, C.c_int]),
        "co_flat_new": (vp, [C.c_int, C.c_int]),
        "co_flat_free": (None, [vp]),
        "co_flat_add": (C.c_int, [vp, C.c_uint32, f32p]),
        "co_flat_add_batch": (C.c_int, [vp, u32p, f32p, C.c_long]),
        "co_flat_remove": (C.c_int, [vp, C.c_uint32]),
        "co_flat_flush": (C.c_int, [vp]),
        "co_flat_size": (C.c_long, [vp]),
        "co_flat_rows": (f32p, [vp]),
        "co_flat_ids": (u32p, [vp]),
        "co_flat_search": (C.c_long, [vp, f32p, C.c_long, C.c_float, u32p, C.c_long, u32p, f32p, i64p]),
        "co_flat_search_batch": (C.c_int, [vp, f32p, C.c_long, C.c_long, C.c_float, C.c_long, u32p, f32p, i64p]),
        "co_ivf_new": (vp, [C.c_int, C.c_int, C.c_int]),
        "co_ivf_free": (None, [vp]),
        "co_ivf_train": (C.c_int, [vp, f32p, C.c_long]),
        "co_ivf_set_centroids": (C.c_int, [vp, f32p]),
        "co_ivf_add": (C.c_int, [vp, C.c_uint32, f32p]),
        "co_ivf_add_batch": (C.c_int, [vp, u32p, f32p, C.c_long]),
        "co_ivf_remove": (C.c_int, [vp, C.c_uint32]),
        "co_ivf_flush": (C.c_int, [vp]),
        "co_ivf_default_nprobes": (C.c_int, [vp]),
        "co_ivf_centroids": (f32p, [vp]),
        "co_ivf_list_len": (C.c_long, [vp, C.c_int]),
        "co_ivf_list_get": (None, [vp, C.c_int, u32p, f32p]),
        "co_ivf_search": (C.c_long, [vp, f32p, C.c_long, C.c_int, C.c_float, u32p, C.c_long, u32p, f32p]),
        "co_pq_new": (vp, [C.c_int, C.c_int, C.c_int, C.c_int]),
        "co_pq_free": (None, [vp]),
        "co_pq_train": (C.c_int, [vp, f32p, C.c_long]),
        "co_pq_set_codebooks": (C.c_int, [vp, f32p]),
        "co_pq_add": (C.c_int, [vp, C.c_uint32, f32p]),
        "co_pq_add_batch": (C.c_int, [vp, u32p, f32p, C.c_long]),
        "co_pq_remove": (C.c_int, [vp, C.c_uint32]),
        "co_pq_flush": (C.c_int, [vp]),
        "co_pq_size": (C.c_long, [vp]),
        "co_pq_codebooks": (f32p, [vp]),
        "co_pq_codes": (u8p, [vp]),
        "co_pq_ids": (u32p, [vp]),
        "co_pq_encode": (None, [vp, f32p, u8p]),
        "co_pq_search": (C.c_long, [vp, f32p, C.c_long, C.c_float, u32p, C.c_long, u32p, f32p]),
        "co_ivfpq_new": (vp, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "co_ivfpq_free": (None, [vp]),
        "co_ivfpq_train": (C.c_int, [vp, f32p, C.c_long]),
        "co_ivfpq_set_trained": (C.c_int, [vp, f32p, f32p]),
        "co_ivfpq_add": (C.c_int, [vp, C.c_uint32, f32p]),
        "co_ivfpq_add_batch": (C.c_int, [vp, u32p, f32p, C.c_long]),
        "co_ivfpq_load_codes": (C.c_int, [vp, u32p, u8p, i32p, C.c_long]),
        "co_ivfpq_remove": (C.c_int, [vp, C.c_uint32]),
        "co_ivfpq_flush": (C.c_int, [vp]),
        "co_ivfpq_default_nprobes": (C.c_int, [vp]),
        "co_ivfpq_centroids": (f32p, [vp]),
        "co_ivfpq_codebooks": (f32p, [vp]),
        "co_ivfpq_list_len": (C.c_long, [vp, C.c_int]),
        "co_ivfpq_list_get": (None, [vp, C.c_int, u32p, u8p]),
        "co_ivfpq_search": (C.c_long, [vp, f32p, C.c_long, C.c_int, C.c_float, u32p, C.c_long, u32p, f32p]),
        "co_ivfpq_last_scanned": (C.c_long, []),
        "co_hnsw_new": (vp, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "co_hnsw_free": (None, [vp]),
        "co_hnsw_add": (C.c_int, [vp, C.c_uint32, f32p, C.c_int]),
        "co_hnsw_add_batch": (C.c_int, [vp, u32p, f32p, i32p, C.c_long]),
        "co_hnsw_load_graph": (C.c_int, [vp, C.c_long, u32p, f32p, i32p, C.POINTER(C.c_longlong), u32p, C.c_uint32, C.c_int]),
        "co_hnsw_remove": (C.c_int, [vp, C.c_uint32]),
        "co_hnsw_size": (C.c_long, [vp]),
        "co_hnsw_max_level": (C.c_int, [vp]),
        "co_hnsw_entry_point": (C.c_uint32, [vp]),
        "co_hnsw_ef_search": (C.c_int, [vp]),
        "co_hnsw_export_nodes": (None, [vp, u32p, i32p, f32p]),
        "co_hnsw_edges": (C.c_int, [vp, C.c_long, C.c_int, u32p]),
        "co_hnsw_search": (C.c_long, [vp, f32p, C.c_long, C.c_int, C.c_float, u32p, C.c_long, u32p, f32p]),
        "co_hnsw_last_dist_evals": (C.c_long, []),
        "co_hnsw_last_expansions": (C.c_long, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _check(rc, what=""):
    if rc < 0:
        raise OracleError(rc, what)
    return rc


def set_fma(on):
    lib().co_set_fma(int(on))


def set_threads(n):
    lib().co_set_threads(int(n))


def distance(metric, a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().co_distance(metric, _p(a, f32p), _p(b, f32p), len(a)))


def normalize(v):
    v = _f32(v)
    out = np.empty_like(v)
    _check(lib().co_normalize(_p(v, f32p), _p(out, f32p), len(v)), "normalize")
    return out


def preprocess(metric, v):
    v = _f32(v)
    out = np.empty_like(v)
    _check(lib().co_preprocess(metric, _p(v, f32p), _p(out, f32p), len(v)), "preprocess")
    return out


def norm(v):
    v = _f32(v)
    return float(lib().co_norm(_p(v, f32p), len(v)))


def sanitize_k(k, n):
    return int(lib().co_sanitize_k(k, n))


def aggregate(kind, ids, scores):
    ids, scores = _u32(ids), _f32(scores)
    n = len(ids)
    oi, os_ = np.empty(max(n, 1), np.uint32), np.empty(max(n, 1), np.float32)
    m = lib().co_aggregate(AGG[kind] if isinstance(kind, str) else kind,
                           _p(ids, u32p), _p(scores, f32p), n, _p(oi, u32p), _p(os_, f32p))
    return oi[:m].copy(), os_[:m].copy()


def autocut(y, cutoff):
    y = _f32(y)
    return int(lib().co_autocut(_p(y, f32p), len(y), cutoff))


def kmeans(vectors, k, metric=L2SQ, max_iter=20):
    v = _f32(vectors)
    n, d = v.shape
    kk = min(k, n)
    cent = np.empty((kk, d), np.float32)
    assign = np.empty(n, np.int32)
    used = lib().co_kmeans(_p(v, f32p), n, d, d, k, metric, max_iter, _p(cent, f32p), _p(assign, i32p))
    return cent[:used], assign


def nearest_centroid(v, centroids, metric):
    v, c = _f32(v), _f32(centroids)
    return int(lib().co_nearest_centroid(_p(v, f32p), _p(c, f32p), c.shape[0], c.shape[1], metric))


class _Index:
    _free = None

    def __del__(self):
        if getattr(self, "h", None) and self._free:
            getattr(lib(), self._free)(self.h)
            self.h = None

    def _filter(self, filter_ids):
        if filter_ids is None or len(filter_ids) == 0:
            return None, 0
        f = _u32(filter_ids)
        return f, len(f)


class Flat(_Index):
    _free = "co_flat_free"

    def __init__(self, dim, metric):
        self.dim, self.metric = dim, metric
        self.h = lib().co_flat_new(dim, metric)
        if not self.h:
            raise OracleError(ERR_ARG, "flat_new")

    def add(self, ids, rows):
        """rows is preprocessed IN PLACE (reference F7) when it is a contiguous float32 array."""
        ids = _u32(np.atleast_1d(ids))
        rows = rows if (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous) else _f32(rows)
        rows2 = rows.reshape(len(ids), self.dim)
        _check(lib().co_flat_add_batch(self.h, _p(ids, u32p), _p(rows2, f32p), len(ids)), "flat_add")

    def remove(self, id_):
        _check(lib().co_flat_remove(self.h, int(id_)), "flat_remove")

    def flush(self):
        _check(lib().co_flat_flush(self.h))

    def __len__(self):
        return int(lib().co_flat_size(self.h))

    def rows(self):
        n = len(self)
        if n == 0:
            return np.empty((0, self.dim), np.float32)
        return np.ctypeslib.as_array(lib().co_flat_rows(self.h), shape=(n, self.dim)).copy()

    def ids(self):
        n = len(self)
        if n == 0:
            return np.empty(0, np.uint32)
        return np.ctypeslib.as_array(lib().co_flat_ids(self.h), shape=(n,)).copy()

    def search(self, query, k=10, threshold=0.0, filter_ids=None, with_pos=False):
        q = _f32(query)
        n = len(self)
        cap = max(n, 1)
        oi, os_, op = np.empty(cap, np.uint32), np.empty(cap, np.float32), np.empty(cap, np.int64)
        f, nf = self._filter(filter_ids)
        m = _check(lib().co_flat_search(self.h, _p(q, f32p), k, threshold, _p(f, u32p), nf,
                                        _p(oi, u32p), _p(os_, f32p), _p(op, i64p)), "flat_search")
        if with_pos:
            return oi[:m].copy(), os_[:m].copy(), op[:m].copy()
        return oi[:m].copy(), os_[:m].copy()

    def search_batch(self, queries, k, threshold=0.0):
        q = _f32(queries)
        nq = q.shape[0]
        kcap = sanitize_k(k, len(self))
        oi = np.zeros((nq, max(kcap, 1)), np.uint32)
        os_ = np.zeros((nq, max(kcap, 1)), np.float32)
        cnt = np.zeros(nq, np.int64)
        _check(lib().co_flat_search_batch(self.h, _p(q, f32p), nq, k, threshold, max(kcap, 1),
                                          _p(oi, u32p), _p(os_, f32p), _p(cnt, i64p)))
        return oi, os_, cnt


class IVF(_Index):
    _free = "co_ivf_free"

    def __init__(self, dim, nlist, metric):
        self.dim, self.nlist, self.metric = dim, nlist, metric
        self.h = lib().co_ivf_new(dim, nlist, metric)
        if not self.h:
            raise OracleError(ERR_ARG, "ivf_new")

    def train(self, rows):
        r = _f32(rows)
        _check(lib().co_ivf_train(self.h, _p(r, f32p), r.shape[0]), "ivf_train")

    def set_centroids(self, c):
        c = _f32(c)
        assert c.shape == (self.nlist, self.dim)
        _check(lib().co_ivf_set_centroids(self.h, _p(c, f32p)))

    def add(self, ids, rows):
        ids = _u32(np.atleast_1d(ids))
        rows = rows if (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous) else _f32(rows)
        _check(lib().co_ivf_add_batch(self.h, _p(ids, u32p), _p(rows.reshape(len(ids), self.dim), f32p), len(ids)), "ivf_add")

    def remove(self, id_):
        _check(lib().co_ivf_remove(self.h, int(id_)), "ivf_remove")

    def flush(self):
        _check(lib().co_ivf_flush(self.h))

    def default_nprobes(self):
        return int(lib().co_ivf_default_nprobes(self.h))

    def centroids(self):
        return np.ctypeslib.as_array(lib().co_ivf_centroids(self.h), shape=(self.nlist, self.dim)).copy()

    def lists(self):
        out = []
        for l in range(self.nlist):
            n = int(lib().co_ivf_list_len(self.h, l))
            ids = np.empty(n, np.uint32)
            rows = np.empty((n, self.dim), np.float32)
            if n:
                lib().co_ivf_list_get(self.h, l, _p(ids, u32p), _p(rows, f32p))
            out.append((ids, rows))
        return out

    def total(self):
        return sum(int(lib().co_ivf_list_len(self.h, l)) for l in range(self.nlist))

    def search(self, query, k=10, nprobes=None, threshold=0.0, filter_ids=None):
        q = _f32(query)
        if nprobes is None:
            nprobes = self.default_nprobes()
        cap = max(self.total(), 1)
        oi, os_ = np.empty(cap, np.uint32), np.empty(cap, np.float32)
        f, nf = self._filter(filter_ids)
        m = _check(lib().co_ivf_search(self.h, _p(q, f32p), k, nprobes, threshold, _p(f, u32p), nf,
                                       _p(oi, u32p), _p(os_, f32p)), "ivf_search")
        return oi[:m].copy(), os_[:m].copy()


class PQ(_Index):
    _free = "co_pq_free"

    def __init__(self, dim, metric, M, nbits):
        self.dim, self.metric, self.M, self.nbits = dim, metric, M, nbits
        self.h = lib().co_pq_new(dim, metric, M, nbits)
        if not self.h:
            raise OracleError(ERR_ARG, "pq_new")
        self.ksub, self.dsub = 1 << nbits, dim // M

    def train(self, rows):
        r = _f32(rows)
        _check(lib().co_pq_train(self.h, _p(r, f32p), r.shape[0]), "pq_train")

    def set_codebooks(self, cb):
        cb = _f32(cb)
        assert cb.size == self.M * self.ksub * self.dsub
        _check(lib().co_pq_set_codebooks(self.h, _p(cb, f32p)))

    def add(self, ids, rows):
        ids = _u32(np.atleast_1d(ids))
        rows = rows if (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous) else _f32(rows)
        _check(lib().co_pq_add_batch(self.h, _p(ids, u32p), _p(rows.reshape(len(ids), self.dim), f32p), len(ids)), "pq_add")

    def remove(self, id_):
        _check(lib().co_pq_remove(self.h, int(id_)), "pq_remove")

    def flush(self):
        _check(lib().co_pq_flush(self.h))

    def __len__(self):
        return int(lib().co_pq_size(self.h))

    def codebooks(self):
        return np.ctypeslib.as_array(lib().co_pq_codebooks(self.h), shape=(self.M, self.ksub, self.dsub)).copy()

    def codes(self):
        n = len(self)
        if n == 0:
            return np.empty((0, self.M), np.uint8)
        return np.ctypeslib.as_array(lib().co_pq_codes(self.h), shape=(n, self.M)).copy()

    def ids(self):
        n = len(self)
        if n == 0:
            return np.empty(0, np.uint32)
        return np.ctypeslib.as_array(lib().co_pq_ids(self.h), shape=(n,)).copy()

    def encode(self, v):
        v = _f32(v)
        code = np.empty(self.M, np.uint8)
        lib().co_pq_encode(self.h, _p(v, f32p), _p(code, u8p))
        return code

    def search(self, query, k=10, threshold=0.0, filter_ids=None):
        q = _f32(query)
        cap = max(len(self), 1)
        oi, os_ = np.empty(cap, np.uint32), np.empty(cap, np.float32)
        f, nf = self._filter(filter_ids)
        m = _check(lib().co_pq_search(self.h, _p(q, f32p), k, threshold, _p(f, u32p), nf,
                                      _p(oi, u32p), _p(os_, f32p)), "pq_search")
        return oi[:m].copy(), os_[:m].copy()


class IVFPQ(_Index):
    _free = "co_ivfpq_free"

    def __init__(self, dim, metric, nlist, M, nbits):
        self.dim, self.metric, self.nlist, self.M, self.nbits = dim, metric, nlist, M, nbits
        self.h = lib().co_ivfpq_new(dim, metric, nlist, M, nbits)
        if not self.h:
            raise OracleError(ERR_ARG, "ivfpq_new")
        self.ksub, self.dsub = 1 << nbits, dim // M

    def train(self, rows):
        r = _f32(rows)
        _check(lib().co_ivfpq_train(self.h, _p(r, f32p), r.shape[0]), "ivfpq_train")

    def set_trained(self, centroids, codebooks):
        c, cb = _f32(centroids), _f32(codebooks)
        _check(lib().co_ivfpq_set_trained(self.h, _p(c, f32p), _p(cb, f32p)))

    def load_codes(self, ids, codes, list_of):
        """Restore path (IVFPQIndex.ReadFrom): stored codes go back to their lists without re-encoding."""
        ids = _u32(np.atleast_1d(ids))
        c = np.ascontiguousarray(codes, dtype=np.uint8).reshape(len(ids), -1)
        lo = np.ascontiguousarray(list_of, dtype=np.int32)
        _check(lib().co_ivfpq_load_codes(self.h, _p(ids, u32p), _p(c, u8p), _p(lo, i32p), len(ids)), "ivfpq_load_codes")

    def add(self, ids, rows):
        ids = _u32(np.atleast_1d(ids))
        rows = rows if (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous) else _f32(rows)
        _check(lib().co_ivfpq_add_batch(self.h, _p(ids, u32p), _p(rows.reshape(len(ids), self.dim), f32p), len(ids)), "ivfpq_add")

    def remove(self, id_):
        _check(lib().co_ivfpq_remove(self.h, int(id_)), "ivfpq_remove")

    def flush(self):
        _check(lib().co_ivfpq_flush(self.h))

    def default_nprobes(self):
        return int(lib().co_ivfpq_default_nprobes(self.h))

    def centroids(self):
        return np.ctypeslib.as_array(lib().co_ivfpq_centroids(self.h), shape=(self.nlist, self.dim)).copy()

    def codebooks(self):
        return np.ctypeslib.as_array(lib().co_ivfpq_codebooks(self.h), shape=(self.M, self.ksub, self.dsub)).copy()

    def lists(self):
        out = []
        for l in range(self.nlist):
            n = int(lib().co_ivfpq_list_len(self.h, l))
            ids = np.empty(n, np.uint32)
            codes = np.empty((n, self.M), np.uint8)
            if n:
                lib().co_ivfpq_list_get(self.h, l, _p(ids, u32p), _p(codes, u8p))
            out.append((ids, codes))
        return out

    def total(self):
        return sum(int(lib().co_ivfpq_list_len(self.h, l)) for l in range(self.nlist))

    def search(self, query, k=10, nprobes=None, threshold=0.0, filter_ids=None):
        q = _f32(query)
        if nprobes is None:
            nprobes = self.default_nprobes()
        cap = max(self.total(), 1)
        oi, os_ = np.empty(cap, np.uint32), np.empty(cap, np.float32)
        f, nf = self._filter(filter_ids)
        m = _check(lib().co_ivfpq_search(self.h, _p(q, f32p), k, nprobes, threshold, _p(f, u32p), nf,
                                         _p(oi, u32p), _p(os_, f32p)), "ivfpq_search")
        return oi[:m].copy(), os_[:m].copy()

    @staticmethod
    def last_scanned():
        return int(lib().co_ivfpq_last_scanned())


def hnsw_random_levels(n, m, seed):
    """hnsw_index.go:474-484 randomLevel with an injected, seeded RNG (the reference draws from
    the unseeded global math/rand/v2): geometric with p = 1/M, capped at 16."""
    rng = np.random.default_rng(seed)
    p = 1.0 / float(m)
    levels = np.zeros(n, np.int32)
    for i in range(n):
        lvl = 0
        while lvl < 16 and rng.random() < p:
            lvl += 1
        levels[i] = lvl
    return levels


class HNSW(_Index):
    _free = "co_hnsw_free"

    def __init__(self, dim, metric, m=16, ef_construction=200, ef_search=200):
        self.dim, self.metric = dim, metric
        self.h = lib().co_hnsw_new(dim, metric, m, ef_construction, ef_search)
        if not self.h:
            raise OracleError(ERR_ARG, "hnsw_new")
        self.m = m if m > 0 else 16

    def add(self, ids, rows, levels):
        ids = _u32(np.atleast_1d(ids))
        levels = np.ascontiguousarray(np.atleast_1d(levels), dtype=np.int32)
        rows = rows if (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous) else _f32(rows)
        _check(lib().co_hnsw_add_batch(self.h, _p(ids, u32p), _p(rows.reshape(len(ids), self.dim), f32p),
                                       _p(levels, i32p), len(ids)), "hnsw_add")

    def load_graph(self, ids, rows, levels, edge_off, edge_ids, entry_id, max_level):
        """Restore path (HNSWIndex.ReadFrom): stored vectors, levels, per-(slot, layer) edge lists, entry point."""
        ids = _u32(np.atleast_1d(ids))
        rows = rows if (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous) else _f32(rows)
        levels = np.ascontiguousarray(levels, dtype=np.int32)
        eo = np.ascontiguousarray(edge_off, dtype=np.int64)
        ei = _u32(edge_ids) if len(edge_ids) else np.zeros(1, np.uint32)
        _check(lib().co_hnsw_load_graph(self.h, len(ids), _p(ids, u32p), _p(rows.reshape(len(ids), self.dim), f32p),
                                        _p(levels, i32p), _p(eo, C.POINTER(C.c_longlong)), _p(ei, u32p), int(entry_id),
                                        int(max_level)), "hnsw_load_graph")

    def remove(self, id_):
        _check(lib().co_hnsw_remove(self.h, int(id_)), "hnsw_remove")

    def __len__(self):
        return int(lib().co_hnsw_size(self.h))

    @property
    def max_level(self):
        return int(lib().co_hnsw_max_level(self.h))

    @property
    def entry_point(self):
        return int(lib().co_hnsw_entry_point(self.h))

    @property
    def ef_search(self):
        return int(lib().co_hnsw_ef_search(self.h))

    def export(self):
        """-> ids[n], levels[n], rows[n,dim], and per layer a CSR (offsets[n+1], neighbour IDs)."""
        n = len(self)
        ids, levels = np.empty(n, np.uint32), np.empty(n, np.int32)
        rows = np.empty((n, self.dim), np.float32)
        lib().co_hnsw_export_nodes(self.h, _p(ids, u32p), _p(levels, i32p), _p(rows, f32p))
        layers = []
        buf = np.empty(4 * self.m + 64, np.uint32)
        for layer in range(self.max_level + 1):
            offs = np.zeros(n + 1, np.int64)
            chunks = []
            for s in range(n):
                c = 0
                if levels[s] >= layer:
                    c = int(lib().co_hnsw_edges(self.h, s, layer, _p(buf, u32p)))
                    chunks.append(buf[:c].copy())
                offs[s + 1] = offs[s] + c
            nbrs = np.concatenate(chunks) if chunks else np.empty(0, np.uint32)
            layers.append((offs, nbrs.astype(np.uint32)))
        return ids, levels, rows, layers

    def search(self, query, k=10, ef_search=0, threshold=0.0, filter_ids=None):
        q = _f32(query)
        cap = max(len(self), 1)
        oi, os_ = np.empty(cap, np.uint32), np.empty(cap, np.float32)
        f, nf = self._filter(filter_ids)
        m = _check(lib().co_hnsw_search(self.h, _p(q, f32p), k, ef_search, threshold, _p(f, u32p), nf,
                                        _p(oi, u32p), _p(os_, f32p)), "hnsw_search")
        return oi[:m].copy(), os_[:m].copy()

    @staticmethod
    def last_counters():
        return int(lib().co_hnsw_last_dist_evals()), int(lib().co_hnsw_last_expansions())
