"""ctypes binding of include/comet_b200.h.  No compute happens in Python."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libcomet_b200.so")

L2, L2SQ, COSINE = 0, 1, 2
METRICS = {"l2": L2, "l2_squared": L2SQ, "cosine": COSINE}
PATH_AUTO, PATH_EXACT, PATH_TENSOR = 0, 1, 2
ROUND_SEPARATE, ROUND_FMA = 0, 1

OK, ERR_INVALID_ARG, ERR_DIM_MISMATCH, ERR_ZERO_VECTOR, ERR_NOT_TRAINED, ERR_NOT_FOUND, ERR_CUDA, \
    ERR_UNSUPPORTED, ERR_TOO_FEW, ERR_BUFFER_TOO_SMALL = range(10)


class CometError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.msg = msg


class SearchParams(C.Structure):
    _fields_ = [("k", C.c_int64), ("threshold", C.c_float), ("nprobes", C.c_int32), ("ef_search", C.c_int32),
                ("filter_ids", C.POINTER(C.c_uint32)), ("nfilter", C.c_int64), ("path", C.c_int32),
                ("reserved", C.c_int32)]


class FlatStats(C.Structure):
    _fields_ = [("path_used", C.c_int32), ("passes", C.c_int32), ("candidates", C.c_int64),
                ("fallback_queries", C.c_int64), ("kernel_launches", C.c_int64)]


f32p, u32p, u8p, i32p, i64p, vp = (C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_void_p)

# name -> (restype, argtypes).  Mirrors include/comet_b200.h one to one (tests check that).
SIGNATURES = {
    "cm_init": (C.c_int, [i32p, C.c_int]),
    "cm_shutdown": (None, []),
    "cm_last_error": (C.c_char_p, []),
    "cm_device_count": (C.c_int, []),
    "cm_set_rounding": (C.c_int, [C.c_int]),
    "cm_get_rounding": (C.c_int, []),
    "cm_version": (C.c_char_p, []),
    "cm_kernel_launches": (C.c_int64, []),
    "cm_profile_enable": (C.c_int, [C.c_int]),
    "cm_profile_reset": (C.c_int, []),
    "cm_profile_get": (C.c_int, [C.c_int, C.POINTER(C.c_double), i64p]),
    "cm_host_alloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
    "cm_host_free": (C.c_int, [vp]),
    "cm_distance_pairs": (C.c_int, [C.c_int, f32p, f32p, C.c_int64, C.c_int, f32p]),
    "cm_preprocess_rows": (C.c_int, [C.c_int, f32p, C.c_int64, C.c_int, i64p]),
    "cm_flat_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(vp)]),
    "cm_flat_destroy": (C.c_int, [vp]),
    "cm_flat_reserve": (C.c_int, [vp, C.c_int64]),
    "cm_flat_add": (C.c_int, [vp, u32p, f32p, C.c_int64, C.c_int]),
    "cm_flat_load_rows": (C.c_int, [vp, u32p, f32p, C.c_int64]),
    "cm_flat_add_device": (C.c_int, [vp, u32p, vp, C.c_int64, vp]),
    "cm_flat_remove": (C.c_int, [vp, C.c_uint32]),
    "cm_flat_flush": (C.c_int, [vp]),
    "cm_flat_size": (C.c_int64, [vp]),
    "cm_flat_dim": (C.c_int, [vp]),
    "cm_flat_metric": (C.c_int, [vp]),
    "cm_flat_get_vector": (C.c_int, [vp, C.c_uint32, f32p]),
    "cm_flat_get_rows": (C.c_int, [vp, i64p, C.c_int64, f32p]),
    "cm_flat_search": (C.c_int, [vp, f32p, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, u32p, f32p,
                                 i64p, i64p]),
    "cm_flat_search_device": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, vp, vp,
                                        vp, vp, vp]),
    "cm_flat_sharded_create": (C.c_int, [C.c_int, C.c_int, i32p, C.c_int, C.c_int64, C.POINTER(vp)]),
    "cm_flat_sharded_destroy": (C.c_int, [vp]),
    "cm_flat_sharded_shards": (C.c_int, [vp]),
    "cm_flat_sharded_size": (C.c_int64, [vp]),
    "cm_flat_sharded_shard_size": (C.c_int, [vp, C.c_int, i64p]),
    "cm_flat_sharded_reserve": (C.c_int, [vp, C.c_int64]),
    "cm_flat_sharded_add": (C.c_int, [vp, u32p, f32p, C.c_int64, C.c_int]),
    "cm_flat_sharded_add_device": (C.c_int, [vp, C.c_int, u32p, vp, C.c_int64, vp]),
    "cm_flat_sharded_remove": (C.c_int, [vp, C.c_uint32]),
    "cm_flat_sharded_flush": (C.c_int, [vp]),
    "cm_flat_sharded_search": (C.c_int, [vp, f32p, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, u32p, f32p,
                                         i64p]),
    "cm_flat_sharded_search_device": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, vp,
                                                vp, vp, vp]),
    "cm_flat_sharded_last_exchange_bytes": (C.c_int64, [vp]),
    "cm_flat_sharded_last_timing": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "cm_flat_sharded_last_stats": (C.c_int, [vp, C.POINTER(FlatStats)]),
    "cm_flat_batcher_create": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(vp)]),
    "cm_flat_batcher_destroy": (C.c_int, [vp]),
    "cm_flat_batcher_search": (C.c_int, [vp, f32p, C.c_int, C.c_int64, C.c_float, C.c_int64, u32p, f32p, i64p]),
    "cm_flat_batcher_stats": (C.c_int, [vp, i64p, i64p]),
    "cm_flat_last_stats": (C.c_int, [vp, C.POINTER(FlatStats)]),
    "cm_ivf_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "cm_ivf_destroy": (C.c_int, [vp]),
    "cm_ivf_set_centroids": (C.c_int, [vp, f32p]),
    "cm_ivf_train": (C.c_int, [vp, f32p, C.c_int64]),
    "cm_ivf_get_centroids": (C.c_int, [vp, f32p]),
    "cm_ivf_trained": (C.c_int, [vp]),
    "cm_ivf_size": (C.c_int64, [vp]),
    "cm_ivf_default_nprobes": (C.c_int, [vp]),
    "cm_ivf_add": (C.c_int, [vp, u32p, f32p, C.c_int64, C.c_int, i32p]),
    "cm_ivf_load_lists": (C.c_int, [vp, u32p, f32p, i32p, C.c_int64]),
    "cm_ivf_last_scanned": (C.c_int64, [vp]),
    "cm_ivf_remove": (C.c_int, [vp, C.c_uint32]),
    "cm_ivf_flush": (C.c_int, [vp]),
    "cm_ivf_get_rows": (C.c_int, [vp, i64p, C.c_int64, f32p]),
    "cm_ivf_search": (C.c_int, [vp, f32p, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, u32p, f32p,
                                i64p, i64p]),
    "cm_ivf_search_device": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, vp, vp,
                                       vp, vp, vp]),
    "cm_ivf_sharded_create": (C.c_int, [C.c_int, C.c_int, C.c_int, i32p, C.c_int, C.POINTER(vp)]),
    "cm_ivf_sharded_destroy": (C.c_int, [vp]),
    "cm_ivf_sharded_shards": (C.c_int, [vp]),
    "cm_ivf_sharded_set_centroids": (C.c_int, [vp, f32p]),
    "cm_ivf_sharded_train": (C.c_int, [vp, f32p, C.c_int64]),
    "cm_ivf_sharded_get_centroids": (C.c_int, [vp, f32p]),
    "cm_ivf_sharded_trained": (C.c_int, [vp]),
    "cm_ivf_sharded_size": (C.c_int64, [vp]),
    "cm_ivf_sharded_shard_size": (C.c_int, [vp, C.c_int, i64p]),
    "cm_ivf_sharded_owner": (C.c_int, [vp, C.c_int]),
    "cm_ivf_sharded_default_nprobes": (C.c_int, [vp]),
    "cm_ivf_sharded_add": (C.c_int, [vp, u32p, f32p, C.c_int64, C.c_int, i32p]),
    "cm_ivf_sharded_remove": (C.c_int, [vp, C.c_uint32]),
    "cm_ivf_sharded_flush": (C.c_int, [vp]),
    "cm_ivf_sharded_rebalance": (C.c_int, [vp]),
    "cm_ivf_sharded_search": (C.c_int, [vp, f32p, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, u32p, f32p, i64p]),
    "cm_ivf_sharded_search_device": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, vp, vp, vp, vp]),
    "cm_ivf_sharded_last_scanned": (C.c_int, [vp, i64p]),
    "cm_pq_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "cm_pq_destroy": (C.c_int, [vp]),
    "cm_pq_set_codebooks": (C.c_int, [vp, f32p]),
    "cm_pq_train": (C.c_int, [vp, f32p, C.c_int64]),
    "cm_pq_get_codebooks": (C.c_int, [vp, f32p]),
    "cm_pq_trained": (C.c_int, [vp]),
    "cm_pq_size": (C.c_int64, [vp]),
    "cm_pq_add": (C.c_int, [vp, u32p, f32p, C.c_int64, C.c_int]),
    "cm_pq_get_codes": (C.c_int, [vp, C.c_int64, C.c_int64, u8p]),
    "cm_pq_load_codes": (C.c_int, [vp, u32p, u8p, C.c_int64]),
    "cm_pq_remove": (C.c_int, [vp, C.c_uint32]),
    "cm_pq_flush": (C.c_int, [vp]),
    "cm_pq_search": (C.c_int, [vp, f32p, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, u32p, f32p, i64p, i64p]),
    "cm_pq_search_device": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, vp, vp, vp, vp, vp]),
    "cm_ivfpq_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "cm_ivfpq_destroy": (C.c_int, [vp]),
    "cm_ivfpq_set_trained": (C.c_int, [vp, f32p, f32p]),
    "cm_ivfpq_train": (C.c_int, [vp, f32p, C.c_int64]),
    "cm_ivfpq_get_trained": (C.c_int, [vp, f32p, f32p]),
    "cm_ivfpq_trained": (C.c_int, [vp]),
    "cm_ivfpq_size": (C.c_int64, [vp]),
    "cm_ivfpq_default_nprobes": (C.c_int, [vp]),
    "cm_ivfpq_add": (C.c_int, [vp, u32p, f32p, C.c_int64, C.c_int, i32p]),
    "cm_ivfpq_get_codes": (C.c_int, [vp, C.c_int64, C.c_int64, u8p]),
    "cm_ivfpq_load_codes": (C.c_int, [vp, u32p, u8p, i32p, C.c_int64]),
    "cm_ivfpq_last_scanned": (C.c_int64, [vp]),
    "cm_ivfpq_remove": (C.c_int, [vp, C.c_uint32]),
    "cm_ivfpq_flush": (C.c_int, [vp]),
    "cm_ivfpq_search": (C.c_int, [vp, f32p, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, u32p, f32p, i64p, i64p]),
    "cm_ivfpq_search_device": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, vp, vp, vp, vp, vp]),
    "cm_pq_sharded_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_int, C.c_int64, C.POINTER(vp)]),
    "cm_pq_sharded_destroy": (C.c_int, [vp]),
    "cm_pq_sharded_shards": (C.c_int, [vp]),
    "cm_pq_sharded_size": (C.c_int64, [vp]),
    "cm_pq_sharded_trained": (C.c_int, [vp]),
    "cm_pq_sharded_train": (C.c_int, [vp, f32p, C.c_int64]),
    "cm_pq_sharded_set_codebooks": (C.c_int, [vp, f32p]),
    "cm_pq_sharded_get_codebooks": (C.c_int, [vp, f32p]),
    "cm_pq_sharded_add": (C.c_int, [vp, u32p, f32p, C.c_int64, C.c_int]),
    "cm_pq_sharded_remove": (C.c_int, [vp, C.c_uint32]),
    "cm_pq_sharded_flush": (C.c_int, [vp]),
    "cm_pq_sharded_search": (C.c_int, [vp, f32p, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, u32p, f32p, i64p]),
    "cm_pq_sharded_search_device": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, vp, vp, vp, vp]),
    "cm_ivfpq_sharded_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_int, C.POINTER(vp)]),
    "cm_ivfpq_sharded_destroy": (C.c_int, [vp]),
    "cm_ivfpq_sharded_shards": (C.c_int, [vp]),
    "cm_ivfpq_sharded_set_trained": (C.c_int, [vp, f32p, f32p]),
    "cm_ivfpq_sharded_train": (C.c_int, [vp, f32p, C.c_int64]),
    "cm_ivfpq_sharded_get_trained": (C.c_int, [vp, f32p, f32p]),
    "cm_ivfpq_sharded_trained": (C.c_int, [vp]),
    "cm_ivfpq_sharded_size": (C.c_int64, [vp]),
    "cm_ivfpq_sharded_shard_size": (C.c_int, [vp, C.c_int, i64p]),
    "cm_ivfpq_sharded_owner": (C.c_int, [vp, C.c_int]),
    "cm_ivfpq_sharded_default_nprobes": (C.c_int, [vp]),
    "cm_ivfpq_sharded_add": (C.c_int, [vp, u32p, f32p, C.c_int64, C.c_int, i32p]),
    "cm_ivfpq_sharded_remove": (C.c_int, [vp, C.c_uint32]),
    "cm_ivfpq_sharded_flush": (C.c_int, [vp]),
    "cm_ivfpq_sharded_rebalance": (C.c_int, [vp]),
    "cm_ivfpq_sharded_search": (C.c_int, [vp, f32p, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, u32p, f32p, i64p]),
    "cm_ivfpq_sharded_search_device": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, vp, vp, vp, vp]),
    "cm_ivfpq_sharded_last_scanned": (C.c_int, [vp, i64p]),
    "cm_hnsw_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "cm_hnsw_destroy": (C.c_int, [vp]),
    "cm_hnsw_size": (C.c_int64, [vp]),
    "cm_hnsw_ef_search": (C.c_int, [vp]),
    "cm_hnsw_load_graph": (C.c_int, [vp, C.c_int64, u32p, f32p, i32p, i64p, u32p, C.c_uint32, C.c_int]),
    "cm_hnsw_add": (C.c_int, [vp, u32p, f32p, i32p, C.c_int64, C.c_int]),
    "cm_hnsw_max_level": (C.c_int, [vp]),
    "cm_hnsw_edge_count": (C.c_int64, [vp]),
    "cm_hnsw_export_graph": (C.c_int, [vp, i32p, i64p, u32p, u32p, i32p]),
    "cm_hnsw_remove": (C.c_int, [vp, C.c_uint32]),
    "cm_hnsw_search": (C.c_int, [vp, f32p, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, u32p, f32p,
                                 i64p, i64p, i64p]),
    "cm_hnsw_search_device": (C.c_int, [vp, vp, C.c_int64, C.c_int, C.POINTER(SearchParams), C.c_int64, vp, vp,
                                        vp, vp, vp, vp]),
    "cm_flat_get_ids": (C.c_int, [vp, C.c_int64, C.c_int64, u32p]),
    "cm_hnsw_flush": (C.c_int, [vp]),
    "cm_ivf_get_ids": (C.c_int, [vp, C.c_int64, C.c_int64, u32p]),
    "cm_pq_get_ids": (C.c_int, [vp, C.c_int64, C.c_int64, u32p]),
    "cm_ivfpq_get_ids": (C.c_int, [vp, C.c_int64, C.c_int64, u32p]),
    "cm_hnsw_get_nodes": (C.c_int, [vp, C.c_int64, C.c_int64, u32p, f32p]),
    "cm_debug_decode_roaring": (C.c_int, [u8p, C.c_int64, u32p, C.c_int64, i64p]),
    "cm_flat_save": (C.c_int, [vp, u8p, C.c_int64, i64p]),
    "cm_flat_load": (C.c_int, [vp, u8p, C.c_int64, i64p]),
    "cm_flat_save_file": (C.c_int, [vp, C.c_char_p]),
    "cm_flat_load_file": (C.c_int, [vp, C.c_char_p]),
    "cm_ivf_save": (C.c_int, [vp, u8p, C.c_int64, i64p]),
    "cm_ivf_load": (C.c_int, [vp, u8p, C.c_int64, i64p]),
    "cm_ivf_save_file": (C.c_int, [vp, C.c_char_p]),
    "cm_ivf_load_file": (C.c_int, [vp, C.c_char_p]),
    "cm_pq_save": (C.c_int, [vp, u8p, C.c_int64, i64p]),
    "cm_pq_load": (C.c_int, [vp, u8p, C.c_int64, i64p]),
    "cm_pq_save_file": (C.c_int, [vp, C.c_char_p]),
    "cm_pq_load_file": (C.c_int, [vp, C.c_char_p]),
    "cm_ivfpq_save": (C.c_int, [vp, u8p, C.c_int64, i64p]),
    "cm_ivfpq_load": (C.c_int, [vp, u8p, C.c_int64, i64p]),
    "cm_ivfpq_save_file": (C.c_int, [vp, C.c_char_p]),
    "cm_ivfpq_load_file": (C.c_int, [vp, C.c_char_p]),
    "cm_hnsw_save": (C.c_int, [vp, u8p, C.c_int64, i64p]),
    "cm_hnsw_load": (C.c_int, [vp, u8p, C.c_int64, i64p]),
    "cm_hnsw_save_file": (C.c_int, [vp, C.c_char_p]),
    "cm_hnsw_load_file": (C.c_int, [vp, C.c_char_p]),
    "cm_merge_shards_device": (C.c_int, [vp, vp, vp, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, vp, vp, vp, vp]),
}

_lib = None


def lib():
    """Load libcomet_b200.so (building it in-tree if the sources are newer).  Raises if it cannot be
    built or loaded -- there is no fallback implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH) or os.environ.get("COMET_B200_REBUILD"):
        from . import build as _b
        _b.build()
    L = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)   # AttributeError if the .so is stale: loud by design
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


PROF_FLAT_SCAN, PROF_FLAT_GEMM, PROF_RESCORE, PROF_SELECT, PROF_IVF_SCAN, PROF_PQ_SCAN, PROF_HNSW, PROF_COARSE = range(8)


def profile_get(cls):
    ms, n = C.c_double(0), C.c_int64(0)
    check(lib().cm_profile_get(cls, C.byref(ms), C.byref(n)))
    return ms.value, n.value


def check(rc):
    if rc != OK:
        raise CometError(rc, lib().cm_last_error().decode("utf-8", "replace"))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def make_params(k=10, threshold=0.0, nprobes=0, ef_search=0, filter_ids=None, path=PATH_AUTO):
    """Returns (params, keepalive)."""
    f = None
    p = SearchParams()
    p.k = int(k)
    p.threshold = float(threshold)
    p.nprobes = int(nprobes)
    p.ef_search = int(ef_search)
    p.path = int(path)
    if filter_ids is not None and len(filter_ids) > 0:
        f = _u32(filter_ids)
        p.filter_ids = ptr(f, u32p)
        p.nfilter = len(f)
    else:
        p.filter_ids = None
        p.nfilter = 0
    return p, f


def merge_shards_device(ids_ptr, scores_ptr, counts_ptr, world, nq, in_stride, k, out_ids_ptr, out_scores_ptr,
                        out_counts_ptr=0, out_stride=None, stream=0):
    """All arguments are device pointers: gathered [world][nq][in_stride] lists -> global top-k."""
    check(lib().cm_merge_shards_device(vp(ids_ptr), vp(scores_ptr), vp(counts_ptr) if counts_ptr else None, int(world),
                                       int(nq), int(in_stride), int(k), int(out_stride or k), vp(out_ids_ptr),
                                       vp(out_scores_ptr), vp(out_counts_ptr) if out_counts_ptr else None, vp(stream)))


def distance_pairs(metric, a, b):
    a, b = _f32(a), _f32(b)
    if a.ndim == 1:
        a, b = a[None, :], b[None, :]
    n, d = a.shape
    out = np.empty(n, np.float32)
    check(lib().cm_distance_pairs(metric, ptr(a, f32p), ptr(b, f32p), n, d, ptr(out, f32p)))
    return out


def preprocess_rows(metric, rows):
    """In place on a contiguous float32 array; raises CometError(ERR_ZERO_VECTOR)."""
    assert rows.dtype == np.float32 and rows.flags.c_contiguous
    r2 = rows.reshape(-1, rows.shape[-1])
    bad = C.c_int64(-1)
    check(lib().cm_preprocess_rows(metric, ptr(r2, f32p), r2.shape[0], r2.shape[1], C.byref(bad)))
    return rows



def save_bytes(kind, handle):
    """WriteTo of index type `kind` ("flat" | "ivf" | "pq" | "ivfpq" | "hnsw") -> bytes (flushes the index)."""
    L = lib()
    n = C.c_int64(0)
    check(getattr(L, f"cm_{kind}_save")(handle, None, 0, C.byref(n)))
    buf = np.zeros(max(n.value, 1), np.uint8)
    check(getattr(L, f"cm_{kind}_save")(handle, ptr(buf, u8p), n.value, C.byref(n)))
    return buf[:n.value].tobytes()


def load_bytes(kind, handle, data):
    """ReadFrom: replaces the state of the pre-constructed index; returns the number of bytes consumed."""
    buf = np.frombuffer(data, dtype=np.uint8)
    used = C.c_int64(0)
    check(getattr(lib(), f"cm_{kind}_load")(handle, ptr(buf, u8p), len(buf), C.byref(used)))
    return used.value


def save_file(kind, handle, path):
    check(getattr(lib(), f"cm_{kind}_save_file")(handle, str(path).encode()))


def load_file(kind, handle, path):
    check(getattr(lib(), f"cm_{kind}_load_file")(handle, str(path).encode()))


def decode_roaring(blob):
    b = np.frombuffer(blob, dtype=np.uint8)
    n = C.c_int64(0)
    check(lib().cm_debug_decode_roaring(ptr(b, u8p) if len(b) else None, len(b), None, 0, C.byref(n)))
    out = np.zeros(max(n.value, 1), np.uint32)
    check(lib().cm_debug_decode_roaring(ptr(b, u8p) if len(b) else None, len(b), ptr(out, u32p), n.value, C.byref(n)))
    return out[:n.value]


class FlatIndex:
    """Thin owner of a cm_flat handle."""

    def __init__(self, dim, metric):
        self.h = vp()
        check(lib().cm_flat_create(int(dim), int(metric), C.byref(self.h)))
        self.dim, self.metric = dim, metric

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().cm_flat_destroy(self.h)
            self.h = None

    def reserve(self, n):
        check(lib().cm_flat_reserve(self.h, int(n)))

    def add(self, ids, rows, writeback=True):
        ids = _u32(np.atleast_1d(ids))
        if not (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous):
            rows = _f32(rows)
        rows2 = rows.reshape(len(ids), self.dim)
        check(lib().cm_flat_add(self.h, ptr(ids, u32p), ptr(rows2, f32p), len(ids), 1 if writeback else 0))

    def load_rows(self, ids, rows):
        """Restore stored (already preprocessed) rows as they are (FlatIndex.ReadFrom)."""
        ids = _u32(np.atleast_1d(ids))
        rows2 = _f32(rows).reshape(len(ids), self.dim)
        check(lib().cm_flat_load_rows(self.h, ptr(ids, u32p), ptr(rows2, f32p), len(ids)))

    def add_device(self, ids, rows_dev_ptr, n, stream=0):
        ids = _u32(ids)
        check(lib().cm_flat_add_device(self.h, ptr(ids, u32p), vp(rows_dev_ptr), int(n), vp(stream)))

    def remove(self, id_):
        check(lib().cm_flat_remove(self.h, int(id_)))

    def flush(self):
        check(lib().cm_flat_flush(self.h))

    def __len__(self):
        return int(lib().cm_flat_size(self.h))

    def get_ids(self, first=0, n=None):
        n = len(self) - first if n is None else n
        out = np.zeros(max(n, 1), np.uint32)
        check(lib().cm_flat_get_ids(self.h, int(first), int(n), ptr(out, u32p)))
        return out[:n]

    def get_vector(self, id_):
        out = np.empty(self.dim, np.float32)
        check(lib().cm_flat_get_vector(self.h, int(id_), ptr(out, f32p)))
        return out

    def get_rows(self, positions):
        pos = np.ascontiguousarray(positions, dtype=np.int64)
        out = np.empty((len(pos), self.dim), np.float32)
        check(lib().cm_flat_get_rows(self.h, ptr(pos, i64p), len(pos), ptr(out, f32p)))
        return out

    def effective_k(self, k):
        n = len(self)
        return n if (k <= 0 or k > n) else k

    def search(self, queries, k=10, threshold=0.0, filter_ids=None, path=PATH_AUTO, with_pos=False):
        """nq independent searchSingleQuery calls -> (ids[nq,K], scores[nq,K], counts[nq])."""
        q = _f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        nq, d = q.shape
        ke = max(self.effective_k(k), 1)
        ids = np.zeros((nq, ke), np.uint32)
        sc = np.zeros((nq, ke), np.float32)
        pos = np.zeros((nq, ke), np.int64) if with_pos else None
        cnt = np.zeros(nq, np.int64)
        p, keep = make_params(k=k, threshold=threshold, filter_ids=filter_ids, path=path)
        check(lib().cm_flat_search(self.h, ptr(q, f32p), nq, d, C.byref(p), ke, ptr(ids, u32p), ptr(sc, f32p),
                                   ptr(pos, i64p), ptr(cnt, i64p)))
        if with_pos:
            return ids, sc, cnt, pos
        return ids, sc, cnt

    def search_device(self, q_ptr, nq, k, out_ids_ptr, out_scores_ptr, out_counts_ptr, out_stride, stream=0,
                      threshold=0.0, path=PATH_AUTO, out_pos_ptr=0):
        p, keep = make_params(k=k, threshold=threshold, path=path)
        check(lib().cm_flat_search_device(self.h, vp(q_ptr), int(nq), self.dim, C.byref(p), int(out_stride),
                                          vp(out_ids_ptr), vp(out_scores_ptr), vp(out_pos_ptr) if out_pos_ptr else None,
                                          vp(out_counts_ptr), vp(stream)))

    def last_stats(self):
        s = FlatStats()
        check(lib().cm_flat_last_stats(self.h, C.byref(s)))
        return {"path_used": s.path_used, "passes": s.passes, "candidates": s.candidates,
                "fallback_queries": s.fallback_queries, "kernel_launches": s.kernel_launches}


class ShardedFlatIndex:
    """Thin owner of a cm_flat_sharded handle: one process, one shard per entry of `devices`."""

    def __init__(self, dim, metric, devices, rows_per_shard):
        self.h = vp()
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        check(lib().cm_flat_sharded_create(int(dim), int(metric), ptr(dv, i32p), len(dv), int(rows_per_shard),
                                           C.byref(self.h)))
        self.dim, self.metric, self.devices = dim, metric, list(devices)

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().cm_flat_sharded_destroy(self.h)
            self.h = None

    def __len__(self):
        return int(lib().cm_flat_sharded_size(self.h))

    def shard_size(self, r):
        n = C.c_int64(0)
        check(lib().cm_flat_sharded_shard_size(self.h, int(r), C.byref(n)))
        return n.value

    def reserve(self, n):
        check(lib().cm_flat_sharded_reserve(self.h, int(n)))

    def add(self, ids, rows, writeback=True):
        ids = _u32(np.atleast_1d(ids))
        if not (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous):
            rows = _f32(rows)
        rows2 = rows.reshape(len(ids), self.dim)
        check(lib().cm_flat_sharded_add(self.h, ptr(ids, u32p), ptr(rows2, f32p), len(ids), 1 if writeback else 0))

    def add_device(self, shard, ids, rows_dev_ptr, n, stream=0):
        ids = _u32(ids)
        check(lib().cm_flat_sharded_add_device(self.h, int(shard), ptr(ids, u32p), vp(rows_dev_ptr), int(n), vp(stream)))

    def remove(self, id_):
        check(lib().cm_flat_sharded_remove(self.h, int(id_)))

    def flush(self):
        check(lib().cm_flat_sharded_flush(self.h))

    def search(self, queries, k=10, threshold=0.0, filter_ids=None, path=PATH_AUTO):
        q = _f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        nq, d = q.shape
        n = len(self)
        ke = max(n if (k <= 0 or k > n) else k, 1)
        ids = np.zeros((nq, ke), np.uint32)
        sc = np.zeros((nq, ke), np.float32)
        cnt = np.zeros(nq, np.int64)
        p, keep = make_params(k=k, threshold=threshold, filter_ids=filter_ids, path=path)
        check(lib().cm_flat_sharded_search(self.h, ptr(q, f32p), nq, d, C.byref(p), ke, ptr(ids, u32p),
                                           ptr(sc, f32p), ptr(cnt, i64p)))
        return ids, sc, cnt

    def search_device(self, q_ptr, nq, k, out_ids_ptr, out_scores_ptr, out_counts_ptr, out_stride, stream=0,
                      threshold=0.0, path=PATH_AUTO):
        p, keep = make_params(k=k, threshold=threshold, path=path)
        check(lib().cm_flat_sharded_search_device(self.h, vp(q_ptr), int(nq), self.dim, C.byref(p), int(out_stride),
                                                  vp(out_ids_ptr), vp(out_scores_ptr), vp(out_counts_ptr), vp(stream)))

    def exchange_bytes(self):
        return int(lib().cm_flat_sharded_last_exchange_bytes(self.h))

    def last_timing(self):
        a, b, c = C.c_double(0), C.c_double(0), C.c_double(0)
        check(lib().cm_flat_sharded_last_timing(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"search_ms_max": a.value, "gather_ms_max": b.value, "merge_ms": c.value}

    def last_stats(self):
        s = FlatStats()
        check(lib().cm_flat_sharded_last_stats(self.h, C.byref(s)))
        return {"path_used": s.path_used, "passes": s.passes, "candidates": s.candidates,
                "fallback_queries": s.fallback_queries, "kernel_launches": s.kernel_launches}


class IVFIndex:
    """Thin owner of a cm_ivf handle."""

    def __init__(self, dim, nlist, metric):
        self.h = vp()
        check(lib().cm_ivf_create(int(dim), int(nlist), int(metric), C.byref(self.h)))
        self.dim, self.nlist, self.metric = dim, nlist, metric

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().cm_ivf_destroy(self.h)
            self.h = None

    def set_centroids(self, c):
        c = _f32(c).reshape(self.nlist, self.dim)
        check(lib().cm_ivf_set_centroids(self.h, ptr(c, f32p)))

    def train(self, rows):
        r = _f32(rows).reshape(-1, self.dim)
        check(lib().cm_ivf_train(self.h, ptr(r, f32p), r.shape[0]))

    def centroids(self):
        out = np.zeros((self.nlist, self.dim), np.float32)
        check(lib().cm_ivf_get_centroids(self.h, ptr(out, f32p)))
        return out

    def add(self, ids, rows, writeback=True):
        ids = _u32(np.atleast_1d(ids))
        if not (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous):
            rows = _f32(rows)
        rows2 = rows.reshape(len(ids), self.dim)
        lists = np.zeros(len(ids), np.int32)
        check(lib().cm_ivf_add(self.h, ptr(ids, u32p), ptr(rows2, f32p), len(ids), 1 if writeback else 0,
                               ptr(lists, i32p)))
        return lists

    def load_lists(self, ids, rows, list_of):
        ids = _u32(np.atleast_1d(ids))
        rows2 = _f32(rows).reshape(len(ids), self.dim)
        lo = np.ascontiguousarray(list_of, dtype=np.int32)
        check(lib().cm_ivf_load_lists(self.h, ptr(ids, u32p), ptr(rows2, f32p), ptr(lo, i32p), len(ids)))

    def get_rows(self, positions):
        pos = np.ascontiguousarray(positions, dtype=np.int64)
        out = np.empty((len(pos), self.dim), np.float32)
        check(lib().cm_ivf_get_rows(self.h, ptr(pos, i64p), len(pos), ptr(out, f32p)))
        return out

    def remove(self, id_):
        check(lib().cm_ivf_remove(self.h, int(id_)))

    def flush(self):
        check(lib().cm_ivf_flush(self.h))

    def __len__(self):
        return int(lib().cm_ivf_size(self.h))

    def default_nprobes(self):
        return int(lib().cm_ivf_default_nprobes(self.h))

    def search(self, queries, k=10, nprobes=None, threshold=0.0, filter_ids=None, out_stride=None):
        q = _f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        nq, d = q.shape
        if nprobes is None:
            nprobes = self.default_nprobes()
        n = len(self)
        stride = out_stride or max(1, n if (k <= 0 or k > n) else k)
        ids = np.zeros((nq, stride), np.uint32)
        sc = np.zeros((nq, stride), np.float32)
        cnt = np.zeros(nq, np.int64)
        p, keep = make_params(k=k, threshold=threshold, nprobes=nprobes, filter_ids=filter_ids)
        check(lib().cm_ivf_search(self.h, ptr(q, f32p), nq, d, C.byref(p), stride, ptr(ids, u32p), ptr(sc, f32p), None,
                                  ptr(cnt, i64p)))
        return ids, sc, cnt


class ShardedIVFIndex:
    """Thin owner of a cm_ivf_sharded handle: one process, lists spread over `devices`."""

    def __init__(self, dim, nlist, metric, devices):
        self.h = vp()
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        check(lib().cm_ivf_sharded_create(int(dim), int(nlist), int(metric), ptr(dv, i32p), len(dv), C.byref(self.h)))
        self.dim, self.nlist, self.metric, self.devices = dim, nlist, metric, list(devices)

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().cm_ivf_sharded_destroy(self.h)
            self.h = None

    def __len__(self):
        return int(lib().cm_ivf_sharded_size(self.h))

    def shard_size(self, r):
        n = C.c_int64(0)
        check(lib().cm_ivf_sharded_shard_size(self.h, int(r), C.byref(n)))
        return n.value

    def owner(self, l):
        return int(lib().cm_ivf_sharded_owner(self.h, int(l)))

    def set_centroids(self, c):
        c = _f32(c).reshape(self.nlist, self.dim)
        check(lib().cm_ivf_sharded_set_centroids(self.h, ptr(c, f32p)))

    def train(self, rows):
        r = _f32(rows)
        check(lib().cm_ivf_sharded_train(self.h, ptr(r, f32p), len(r)))

    def centroids(self):
        out = np.empty((self.nlist, self.dim), np.float32)
        check(lib().cm_ivf_sharded_get_centroids(self.h, ptr(out, f32p)))
        return out

    def add(self, ids, rows, writeback=True):
        ids = _u32(np.atleast_1d(ids))
        if not (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous):
            rows = _f32(rows)
        rows2 = rows.reshape(len(ids), self.dim)
        lists = np.full(len(ids), -1, np.int32)
        check(lib().cm_ivf_sharded_add(self.h, ptr(ids, u32p), ptr(rows2, f32p), len(ids), 1 if writeback else 0, ptr(lists, i32p)))
        return lists

    def remove(self, id_):
        check(lib().cm_ivf_sharded_remove(self.h, int(id_)))

    def flush(self):
        check(lib().cm_ivf_sharded_flush(self.h))

    def rebalance(self):
        check(lib().cm_ivf_sharded_rebalance(self.h))

    def default_nprobes(self):
        return int(lib().cm_ivf_sharded_default_nprobes(self.h))

    def search(self, queries, k=10, nprobes=None, threshold=0.0, filter_ids=None, out_stride=None):
        q = _f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        nq, d = q.shape
        if nprobes is None:
            nprobes = self.default_nprobes()
        n = len(self)
        stride = out_stride or max(1, n if (k <= 0 or k > n) else k)
        ids = np.zeros((nq, stride), np.uint32)
        sc = np.zeros((nq, stride), np.float32)
        cnt = np.zeros(nq, np.int64)
        p, keep = make_params(k=k, threshold=threshold, nprobes=nprobes, filter_ids=filter_ids)
        check(lib().cm_ivf_sharded_search(self.h, ptr(q, f32p), nq, d, C.byref(p), stride, ptr(ids, u32p), ptr(sc, f32p),
                                          ptr(cnt, i64p)))
        return ids, sc, cnt

    def last_scanned(self):
        out = np.zeros(len(self.devices), np.int64)
        check(lib().cm_ivf_sharded_last_scanned(self.h, ptr(out, i64p)))
        return out


class _ADCIndex:
    """Shared ctypes plumbing of the PQ and IVFPQ handles."""
    _p = ""

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            getattr(lib(), self._p + "_destroy")(self.h)
            self.h = None

    def __len__(self):
        return int(getattr(lib(), self._p + "_size")(self.h))

    def remove(self, id_):
        check(getattr(lib(), self._p + "_remove")(self.h, int(id_)))

    def flush(self):
        check(getattr(lib(), self._p + "_flush")(self.h))

    def codes(self):
        n = len(self)
        out = np.zeros((n, self.M), np.uint8)
        if n:
            check(getattr(lib(), self._p + "_get_codes")(self.h, 0, n, ptr(out, u8p)))
        return out

    def _rows(self, ids, rows):
        ids = _u32(np.atleast_1d(ids))
        if not (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous):
            rows = _f32(rows)
        return ids, rows.reshape(len(ids), self.dim)

    def _search(self, queries, k, threshold, nprobes, filter_ids, out_stride):
        q = _f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        nq, d = q.shape
        n = len(self)
        stride = out_stride or max(1, n if (k <= 0 or k > n) else k)
        ids = np.zeros((nq, stride), np.uint32)
        sc = np.zeros((nq, stride), np.float32)
        cnt = np.zeros(nq, np.int64)
        p, keep = make_params(k=k, threshold=threshold, nprobes=nprobes, filter_ids=filter_ids)
        check(getattr(lib(), self._p + "_search")(self.h, ptr(q, f32p), nq, d, C.byref(p), stride, ptr(ids, u32p),
                                                  ptr(sc, f32p), None, ptr(cnt, i64p)))
        return ids, sc, cnt


class PQIndex(_ADCIndex):
    _p = "cm_pq"

    def __init__(self, dim, metric, M, nbits):
        self.h = vp()
        check(lib().cm_pq_create(int(dim), int(metric), int(M), int(nbits), C.byref(self.h)))
        self.dim, self.metric, self.M, self.nbits = dim, metric, M, nbits

    def set_codebooks(self, cb):
        cb = _f32(cb)
        check(lib().cm_pq_set_codebooks(self.h, ptr(cb, f32p)))

    def train(self, rows):
        r = _f32(rows).reshape(-1, self.dim)
        check(lib().cm_pq_train(self.h, ptr(r, f32p), r.shape[0]))

    def codebooks(self):
        ksub = 1 << self.nbits
        out = np.zeros((self.M, ksub, self.dim // self.M), np.float32)
        check(lib().cm_pq_get_codebooks(self.h, ptr(out, f32p)))
        return out

    def add(self, ids, rows, writeback=True):
        ids, rows2 = self._rows(ids, rows)
        check(lib().cm_pq_add(self.h, ptr(ids, u32p), ptr(rows2, f32p), len(ids), 1 if writeback else 0))

    def load_codes(self, ids, codes):
        ids = _u32(np.atleast_1d(ids))
        c = np.ascontiguousarray(codes, dtype=np.uint8).reshape(len(ids), self.M)
        check(lib().cm_pq_load_codes(self.h, ptr(ids, u32p), ptr(c, u8p), len(ids)))

    def search(self, queries, k=10, threshold=0.0, filter_ids=None, out_stride=None):
        return self._search(queries, k, threshold, 0, filter_ids, out_stride)


class IVFPQIndex(_ADCIndex):
    _p = "cm_ivfpq"

    def __init__(self, dim, metric, nlist, M, nbits):
        self.h = vp()
        check(lib().cm_ivfpq_create(int(dim), int(metric), int(nlist), int(M), int(nbits), C.byref(self.h)))
        self.dim, self.metric, self.nlist, self.M, self.nbits = dim, metric, nlist, M, nbits

    def set_trained(self, centroids, codebooks):
        c, cb = _f32(centroids), _f32(codebooks)
        check(lib().cm_ivfpq_set_trained(self.h, ptr(c, f32p), ptr(cb, f32p)))

    def default_nprobes(self):
        return int(lib().cm_ivfpq_default_nprobes(self.h))

    def train(self, rows):
        r = _f32(rows).reshape(-1, self.dim)
        check(lib().cm_ivfpq_train(self.h, ptr(r, f32p), r.shape[0]))

    def trained_state(self):
        ksub = 1 << self.nbits
        c = np.zeros((self.nlist, self.dim), np.float32)
        cb = np.zeros((self.M, ksub, self.dim // self.M), np.float32)
        check(lib().cm_ivfpq_get_trained(self.h, ptr(c, f32p), ptr(cb, f32p)))
        return c, cb

    def add(self, ids, rows, writeback=True):
        ids, rows2 = self._rows(ids, rows)
        lists = np.zeros(len(ids), np.int32)
        check(lib().cm_ivfpq_add(self.h, ptr(ids, u32p), ptr(rows2, f32p), len(ids), 1 if writeback else 0,
                                 ptr(lists, i32p)))
        return lists

    def load_codes(self, ids, codes, list_of):
        ids = _u32(np.atleast_1d(ids))
        c = np.ascontiguousarray(codes, dtype=np.uint8).reshape(len(ids), self.M)
        lo = np.ascontiguousarray(list_of, dtype=np.int32)
        check(lib().cm_ivfpq_load_codes(self.h, ptr(ids, u32p), ptr(c, u8p), ptr(lo, i32p), len(ids)))

    def search(self, queries, k=10, nprobes=None, threshold=0.0, filter_ids=None, out_stride=None):
        if nprobes is None:
            nprobes = self.default_nprobes()
        return self._search(queries, k, threshold, nprobes, filter_ids, out_stride)


class ShardedPQIndex:
    """Thin owner of a cm_pq_sharded handle: one process, code rows sharded over `devices`."""

    def __init__(self, dim, metric, M, nbits, devices, rows_per_shard):
        self.h = vp()
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        check(lib().cm_pq_sharded_create(int(dim), int(metric), int(M), int(nbits), ptr(dv, i32p), len(dv), int(rows_per_shard),
                                         C.byref(self.h)))
        self.dim, self.metric, self.M, self.nbits, self.devices = dim, metric, M, nbits, list(devices)

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().cm_pq_sharded_destroy(self.h)
            self.h = None

    def __len__(self):
        return int(lib().cm_pq_sharded_size(self.h))

    def train(self, rows):
        r = _f32(rows)
        check(lib().cm_pq_sharded_train(self.h, ptr(r, f32p), len(r)))

    def set_codebooks(self, cb):
        c = _f32(cb)
        check(lib().cm_pq_sharded_set_codebooks(self.h, ptr(c, f32p)))

    def codebooks(self):
        out = np.empty((self.M, 1 << self.nbits, self.dim // self.M), np.float32)
        check(lib().cm_pq_sharded_get_codebooks(self.h, ptr(out, f32p)))
        return out

    def add(self, ids, rows, writeback=True):
        ids = _u32(np.atleast_1d(ids))
        if not (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous):
            rows = _f32(rows)
        rows2 = rows.reshape(len(ids), self.dim)
        check(lib().cm_pq_sharded_add(self.h, ptr(ids, u32p), ptr(rows2, f32p), len(ids), 1 if writeback else 0))

    def remove(self, id_):
        check(lib().cm_pq_sharded_remove(self.h, int(id_)))

    def flush(self):
        check(lib().cm_pq_sharded_flush(self.h))

    def search(self, queries, k=10, threshold=0.0, filter_ids=None):
        q = _f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        nq, d = q.shape
        n = len(self)
        stride = max(1, n if (k <= 0 or k > n) else k)
        ids = np.zeros((nq, stride), np.uint32)
        sc = np.zeros((nq, stride), np.float32)
        cnt = np.zeros(nq, np.int64)
        p, keep = make_params(k=k, threshold=threshold, filter_ids=filter_ids)
        check(lib().cm_pq_sharded_search(self.h, ptr(q, f32p), nq, d, C.byref(p), stride, ptr(ids, u32p), ptr(sc, f32p), ptr(cnt, i64p)))
        return ids, sc, cnt


class ShardedIVFPQIndex:
    """Thin owner of a cm_ivfpq_sharded handle: one process, code lists spread over `devices`."""

    def __init__(self, dim, metric, nlist, M, nbits, devices):
        self.h = vp()
        dv = np.ascontiguousarray(devices, dtype=np.int32)
        check(lib().cm_ivfpq_sharded_create(int(dim), int(metric), int(nlist), int(M), int(nbits), ptr(dv, i32p), len(dv),
                                            C.byref(self.h)))
        self.dim, self.metric, self.nlist, self.M, self.nbits, self.devices = dim, metric, nlist, M, nbits, list(devices)

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().cm_ivfpq_sharded_destroy(self.h)
            self.h = None

    def __len__(self):
        return int(lib().cm_ivfpq_sharded_size(self.h))

    def shard_size(self, r):
        n = C.c_int64(0)
        check(lib().cm_ivfpq_sharded_shard_size(self.h, int(r), C.byref(n)))
        return n.value

    def owner(self, l):
        return int(lib().cm_ivfpq_sharded_owner(self.h, int(l)))

    def set_trained(self, centroids, codebooks):
        c, b = _f32(centroids), _f32(codebooks)
        check(lib().cm_ivfpq_sharded_set_trained(self.h, ptr(c, f32p), ptr(b, f32p)))

    def train(self, rows):
        r = _f32(rows)
        check(lib().cm_ivfpq_sharded_train(self.h, ptr(r, f32p), len(r)))

    def trained_state(self):
        ksub, dsub = 1 << self.nbits, self.dim // self.M
        c = np.empty((self.nlist, self.dim), np.float32)
        b = np.empty((self.M, ksub, dsub), np.float32)
        check(lib().cm_ivfpq_sharded_get_trained(self.h, ptr(c, f32p), ptr(b, f32p)))
        return c, b

    def add(self, ids, rows, writeback=True):
        ids = _u32(np.atleast_1d(ids))
        if not (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous):
            rows = _f32(rows)
        rows2 = rows.reshape(len(ids), self.dim)
        lists = np.full(len(ids), -1, np.int32)
        check(lib().cm_ivfpq_sharded_add(self.h, ptr(ids, u32p), ptr(rows2, f32p), len(ids), 1 if writeback else 0, ptr(lists, i32p)))
        return lists

    def remove(self, id_):
        check(lib().cm_ivfpq_sharded_remove(self.h, int(id_)))

    def flush(self):
        check(lib().cm_ivfpq_sharded_flush(self.h))

    def rebalance(self):
        check(lib().cm_ivfpq_sharded_rebalance(self.h))

    def default_nprobes(self):
        return int(lib().cm_ivfpq_sharded_default_nprobes(self.h))

    def search(self, queries, k=10, nprobes=None, threshold=0.0, filter_ids=None, out_stride=None):
        q = _f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        nq, d = q.shape
        if nprobes is None:
            nprobes = self.default_nprobes()
        n = len(self)
        stride = out_stride or max(1, n if (k <= 0 or k > n) else k)
        ids = np.zeros((nq, stride), np.uint32)
        sc = np.zeros((nq, stride), np.float32)
        cnt = np.zeros(nq, np.int64)
        p, keep = make_params(k=k, threshold=threshold, nprobes=nprobes, filter_ids=filter_ids)
        check(lib().cm_ivfpq_sharded_search(self.h, ptr(q, f32p), nq, d, C.byref(p), stride, ptr(ids, u32p), ptr(sc, f32p),
                                            ptr(cnt, i64p)))
        return ids, sc, cnt

    def last_scanned(self):
        out = np.zeros(len(self.devices), np.int64)
        check(lib().cm_ivfpq_sharded_last_scanned(self.h, ptr(out, i64p)))
        return out


class HNSWIndex:
    """Thin owner of a cm_hnsw handle (search on device; the graph is uploaded)."""

    def __init__(self, dim, metric, m=16, ef_construction=200, ef_search=200):
        self.h = vp()
        check(lib().cm_hnsw_create(int(dim), int(metric), int(m), int(ef_construction), int(ef_search), C.byref(self.h)))
        self.dim, self.metric = dim, metric

    def __del__(self):
        if getattr(self, "h", None) and lib is not None:
            lib().cm_hnsw_destroy(self.h)
            self.h = None

    def __len__(self):
        return int(lib().cm_hnsw_size(self.h))

    def ef_search(self):
        return int(lib().cm_hnsw_ef_search(self.h))

    def load_graph(self, ids, rows, levels, layers, entry_id, max_level):
        """layers: per layer (offsets[n+1], neighbour ids) as exported by the builder."""
        ids = _u32(ids)
        rows = _f32(rows).reshape(len(ids), self.dim)
        levels = np.ascontiguousarray(levels, dtype=np.int32)
        n = len(ids)
        offs = [0]
        chunks = []
        for s in range(n):
            for layer in range(int(levels[s]) + 1):
                lo, nb = layers[layer]
                seg = nb[lo[s]:lo[s + 1]]
                chunks.append(seg)
                offs.append(offs[-1] + len(seg))
        edge_off = np.asarray(offs, dtype=np.int64)
        edge_ids = _u32(np.concatenate(chunks)) if chunks and offs[-1] > 0 else np.zeros(1, np.uint32)
        check(lib().cm_hnsw_load_graph(self.h, n, ptr(ids, u32p), ptr(rows, f32p), ptr(levels, i32p),
                                       ptr(edge_off, i64p), ptr(edge_ids, u32p), int(entry_id), int(max_level)))

    def add(self, ids, rows, levels, writeback=True):
        """n successive HNSWIndex.Add calls on the device with the caller's level draws."""
        ids = _u32(np.atleast_1d(ids))
        if not (isinstance(rows, np.ndarray) and rows.dtype == np.float32 and rows.flags.c_contiguous):
            rows = _f32(rows)
        rows2 = rows.reshape(len(ids), self.dim)
        lv = np.ascontiguousarray(np.atleast_1d(levels), dtype=np.int32)
        check(lib().cm_hnsw_add(self.h, ptr(ids, u32p), ptr(rows2, f32p), ptr(lv, i32p), len(ids), 1 if writeback else 0))

    def max_level(self):
        return int(lib().cm_hnsw_max_level(self.h))

    def export_graph(self):
        """-> levels[n], edge_off[pairs+1], edge_ids, entry_id, max_level (pairs ordered by (slot, layer))."""
        n = len(self)
        ne = int(lib().cm_hnsw_edge_count(self.h))
        levels = np.zeros(max(n, 1), np.int32)
        # first call with a generous offsets buffer: at most 17 layers per node
        edge_off = np.zeros(n * 17 + 1, np.int64)
        edge_ids = np.zeros(max(ne, 1), np.uint32)
        entry, ml = C.c_uint32(0), C.c_int(0)
        check(lib().cm_hnsw_export_graph(self.h, ptr(levels, i32p), ptr(edge_off, i64p), ptr(edge_ids, u32p), C.byref(entry),
                                         C.byref(ml)))
        pairs = int((levels[:n] + 1).sum())
        return levels[:n], edge_off[:pairs + 1], edge_ids[:ne], entry.value, ml.value

    def remove(self, id_):
        check(lib().cm_hnsw_remove(self.h, int(id_)))

    def flush(self):
        check(lib().cm_hnsw_flush(self.h))

    def search(self, queries, k=10, ef_search=0, threshold=0.0, filter_ids=None, with_work=False):
        q = _f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        nq, d = q.shape
        ef = ef_search if ef_search > 0 else self.ef_search()
        stride = max(1, ef if (k <= 0 or k > ef) else k)
        ids = np.zeros((nq, stride), np.uint32)
        sc = np.zeros((nq, stride), np.float32)
        cnt = np.zeros(nq, np.int64)
        work = np.zeros((nq, 2), np.int64)
        p, keep = make_params(k=k, threshold=threshold, ef_search=ef_search, filter_ids=filter_ids)
        check(lib().cm_hnsw_search(self.h, ptr(q, f32p), nq, d, C.byref(p), stride, ptr(ids, u32p), ptr(sc, f32p), None,
                                   ptr(cnt, i64p), ptr(work, i64p)))
        if with_work:
            return ids, sc, cnt, work
        return ids, sc, cnt


class FlatBatcher:
    """Coalesces concurrent single-query searches on a FlatIndex into device batches (cm_flat_batcher_*)."""

    def __init__(self, index, max_batch=512, max_wait_us=200):
        self.index = index          # keeps the index alive
        self.h = vp()
        check(lib().cm_flat_batcher_create(index.h, int(max_batch), int(max_wait_us), C.byref(self.h)))

    def close(self):
        if getattr(self, "h", None):
            lib().cm_flat_batcher_destroy(self.h)
            self.h = None

    __del__ = close

    def search(self, query, k=10, threshold=0.0):
        q = _f32(query).reshape(-1)
        ke = max(self.index.effective_k(k), 1)
        ids = np.zeros(ke, np.uint32)
        sc = np.zeros(ke, np.float32)
        cnt = C.c_int64(0)
        check(lib().cm_flat_batcher_search(self.h, ptr(q, f32p), len(q), int(k), float(threshold), ke, ptr(ids, u32p),
                                           ptr(sc, f32p), C.byref(cnt)))
        return ids[:cnt.value], sc[:cnt.value]

    def stats(self):
        b, r = C.c_int64(0), C.c_int64(0)
        check(lib().cm_flat_batcher_stats(self.h, C.byref(b), C.byref(r)))
        return b.value, r.value
