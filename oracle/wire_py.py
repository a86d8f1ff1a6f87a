"""CPU restatement of the reference's index wire formats (WriteTo / ReadFrom) and of HNSWIndex.Flush.

TEST INFRASTRUCTURE ONLY (like the rest of oracle/): imported by tests/ to check the product's cm_*_save /
cm_*_load byte for byte.  Nothing under comet_b200/ may import this module.

Every writer follows the field order of the reference function it cites; all integers little-endian
(binary.Write(w, binary.LittleEndian, ...)).  States are plain numpy arrays / dicts, so the same state can be
handed to the oracle's C indexes (oracle_py) and to the product (capi).

Parity status: pinned by hand-built streams decoded field by field (tests/test_wire_cpu.py) written from the
reference's format comments (flat_index.go:348-359, ivf_index.go:440-461, pq_index.go:480-502,
ivfpq_index.go:507-537, hnsw_index.go:701-727); the reference holds no golden byte fixtures for these formats
(its tests round-trip through its own reader), and Go is not available here to produce one.

The roaring blob: WriteTo always flushes first, so the deleted set it writes is empty; roaring v1.9.4 (go.mod)
ToBytes of an empty bitmap is the portable format's 8 bytes -- cookie 12346 (no run containers), 0 containers.
`roaring_encode` / `roaring_decode` restate the published RoaringFormatSpec for array / bitmap / run containers
(the reader must accept whatever a bitmap holds).
"""
from __future__ import annotations

import math
import struct

import numpy as np

KIND = {0: "l2", 1: "l2_squared", 2: "cosine"}          # distance.go:21-38
EMPTY_ROARING = struct.pack("<II", 12346, 0)


class WireError(ValueError):
    pass


# ---- roaring portable format (RoaringFormatSpec) ---------------------------------------------------------------
def roaring_encode(ids, use_runs=False):
    """Portable serialisation of a set of uint32 (array / bitmap containers; run containers if use_runs)."""
    ids = sorted(set(int(i) for i in ids))
    groups = {}
    for v in ids:
        groups.setdefault(v >> 16, []).append(v & 0xFFFF)
    keys = sorted(groups)
    n = len(keys)
    if not use_runs:
        out = struct.pack("<II", 12346, n)
    else:
        out = struct.pack("<I", 12347 | ((n - 1) << 16))
        flags = bytearray((n + 7) // 8)
        for i in range(n):
            flags[i // 8] |= 1 << (i % 8)
        out += bytes(flags)
    for k in keys:
        out += struct.pack("<HH", k, len(groups[k]) - 1)
    bodies = []
    for k in keys:
        lows = groups[k]
        if use_runs:
            runs = []
            start = prev = lows[0]
            for v in lows[1:]:
                if v == prev + 1:
                    prev = v
                    continue
                runs.append((start, prev - start))
                start = prev = v
            runs.append((start, prev - start))
            bodies.append(struct.pack("<H", len(runs)) + b"".join(struct.pack("<HH", a, b) for a, b in runs))
        elif len(lows) > 4096:
            words = np.zeros(1024, np.uint64)
            for v in lows:
                words[v >> 6] |= np.uint64(1) << np.uint64(v & 63)
            bodies.append(words.astype("<u8").tobytes())
        else:
            bodies.append(np.asarray(lows, dtype="<u2").tobytes())
    if (not use_runs) or n >= 4:                              # offset header
        at = len(out) + 4 * n
        for b in bodies:
            out += struct.pack("<I", at)
            at += len(b)
    return out + b"".join(bodies)


def roaring_decode(blob):
    if len(blob) == 0:
        return []
    (cookie,) = struct.unpack_from("<I", blob, 0)
    at = 4
    run_flags = None
    if cookie & 0xFFFF == 12347:
        n = (cookie >> 16) + 1
        nb = (n + 7) // 8
        run_flags = blob[at:at + nb]
        at += nb
    elif cookie == 12346:
        (n,) = struct.unpack_from("<I", blob, at)
        at += 4
    else:
        raise WireError(f"unknown roaring cookie {cookie}")
    heads = [struct.unpack_from("<HH", blob, at + 4 * i) for i in range(n)]
    at += 4 * n
    if run_flags is None or n >= 4:
        at += 4 * n
    out = []
    for i, (key, card1) in enumerate(heads):
        hi = key << 16
        card = card1 + 1
        if run_flags is not None and (run_flags[i // 8] >> (i % 8)) & 1:
            (nr,) = struct.unpack_from("<H", blob, at)
            at += 2
            for _ in range(nr):
                start, len1 = struct.unpack_from("<HH", blob, at)
                at += 4
                out.extend(hi | v for v in range(start, start + len1 + 1))
        elif card > 4096:
            words = np.frombuffer(blob, dtype="<u8", count=1024, offset=at)
            at += 8192
            for w, bits in enumerate(words):
                bits = int(bits)
                while bits:
                    b = (bits & -bits).bit_length() - 1
                    out.append(hi | (w * 64 + b))
                    bits &= bits - 1
        else:
            lows = np.frombuffer(blob, dtype="<u2", count=card, offset=at)
            at += 2 * card
            out.extend(hi | int(v) for v in lows)
    return out


# ---- shared pieces -------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, data):
        self.b = memoryview(data)
        self.at = 0

    def take(self, n, what):
        if self.at + n > len(self.b):
            raise WireError(f"failed to read {what}: unexpected EOF")
        v = self.b[self.at:self.at + n]
        self.at += n
        return v

    def u32(self, what):
        return struct.unpack("<I", self.take(4, what))[0]

    def i32(self, what):
        return struct.unpack("<i", self.take(4, what))[0]

    def u8(self, what):
        return self.take(1, what)[0]

    def f64(self, what):
        return struct.unpack("<d", self.take(8, what))[0]

    def f32s(self, n, what):
        return np.frombuffer(self.take(4 * n, what), dtype="<f4").copy()


def _header(magic, dim, metric):
    kind = KIND[metric].encode()
    return magic + struct.pack("<III", 1, dim, len(kind)) + kind


def _read_header(r, magic, dim, metric):
    m = bytes(r.take(4, "magic number"))
    if m != magic:
        raise WireError(f"invalid magic number: expected '{magic.decode()}', got '{m.decode(errors='replace')}'")
    version = r.u32("version")
    if version != 1:
        raise WireError(f"unsupported version: {version}")
    d = r.u32("dimensionality")
    if d != dim:
        raise WireError(f"dimension mismatch: index has dim={dim}, serialized data has dim={d}")
    kind = bytes(r.take(r.u32("distance kind length"), "distance kind")).decode()
    if kind != KIND[metric]:
        raise WireError(f"distance kind mismatch: index uses '{KIND[metric]}', serialized data uses '{kind}'")


def _bitmap(deleted_ids):
    blob = roaring_encode(deleted_ids) if len(deleted_ids) else EMPTY_ROARING
    return struct.pack("<I", len(blob)) + blob


def _read_bitmap(r):
    return roaring_decode(bytes(r.take(r.u32("bitmap size"), "bitmap data")))


def _f32(a):
    return np.ascontiguousarray(a, dtype="<f4")


# ---- FLAT (flat_index.go:366-470 WriteTo, :488-614 ReadFrom) ----------------------------------------------------
def write_flat(dim, metric, ids, rows, deleted_ids=()):
    """ids[n], rows[n, dim]: the STORED vectors in scan order (the caller has flushed, as WriteTo does)."""
    rows = _f32(rows).reshape(len(ids), dim)
    out = [_header(b"FLAT", dim, metric), struct.pack("<I", len(ids))]
    for i, id_ in enumerate(ids):
        out.append(struct.pack("<II", int(id_), dim))
        out.append(rows[i].tobytes())
    out.append(_bitmap(deleted_ids))
    return b"".join(out)


def read_flat(data, dim, metric):
    r = _Reader(data)
    _read_header(r, b"FLAT", dim, metric)
    n = r.u32("vector count")
    ids = np.zeros(n, np.uint32)
    rows = np.zeros((n, dim), np.float32)
    for i in range(n):
        ids[i] = r.u32("vector ID")
        vd = r.u32("vector dimension")
        if vd != dim:
            raise WireError(f"vector {i} has dimension {vd}, expected {dim}")
        rows[i] = r.f32s(dim, "vector data")
    deleted = _read_bitmap(r)
    return {"ids": ids, "rows": rows, "deleted": deleted, "consumed": r.at}


# ---- IVFX (ivf_index.go:468-610 WriteTo, :611-785 ReadFrom) -----------------------------------------------------
def write_ivf(dim, metric, nlist, centroids, lists, deleted_ids=()):
    """centroids: [nlist, dim] or None (untrained); lists: per list a (ids, rows) pair in list order."""
    out = [_header(b"IVFX", dim, metric), struct.pack("<IB", nlist, 1 if centroids is not None else 0)]
    if centroids is not None:
        c = _f32(centroids).reshape(nlist, dim)
        for l in range(nlist):
            out.append(struct.pack("<I", dim))
            out.append(c[l].tobytes())
    out.append(struct.pack("<I", len(lists)))
    for ids, rows in lists:
        rows = _f32(rows).reshape(len(ids), dim)
        out.append(struct.pack("<I", len(ids)))
        for i, id_ in enumerate(ids):
            out.append(struct.pack("<I", int(id_)))
            out.append(rows[i].tobytes())
    out.append(_bitmap(deleted_ids))
    return b"".join(out)


def read_ivf(data, dim, metric, nlist):
    r = _Reader(data)
    _read_header(r, b"IVFX", dim, metric)
    nl = r.u32("nlist")
    if nl != nlist:
        raise WireError(f"nlist mismatch: index has nlist={nlist}, serialized data has nlist={nl}")
    trained = r.u8("trained flag") == 1
    centroids = None
    if trained:
        centroids = np.zeros((nlist, dim), np.float32)
        for l in range(nlist):
            sz = r.u32("centroid size")
            centroids[l, :] = r.f32s(sz, "centroid data")[:dim]
    lists = []
    for _ in range(r.u32("list count")):
        sz = r.u32("list size")
        ids = np.zeros(sz, np.uint32)
        rows = np.zeros((sz, dim), np.float32)
        for i in range(sz):
            ids[i] = r.u32("list vector ID")
            rows[i] = r.f32s(dim, "list vector data")
        lists.append((ids, rows))
    deleted = _read_bitmap(r)
    return {"trained": trained, "centroids": centroids, "lists": lists, "deleted": deleted, "consumed": r.at}


# ---- PQIX (pq_index.go:509-650 WriteTo, :652-846 ReadFrom) ------------------------------------------------------
def _pq_params(M, nbits, ksub, dsub):
    return struct.pack("<IIII", M, nbits, ksub, dsub)


def write_pq(dim, metric, M, nbits, codebooks, ids, codes, deleted_ids=()):
    """codebooks: [M, Ksub, dsub] or None; codes: uint8 [n, M] in arrival order."""
    ksub, dsub = 1 << nbits, dim // M
    out = [_header(b"PQIX", dim, metric), _pq_params(M, nbits, ksub, dsub), struct.pack("<B", 1 if codebooks is not None else 0)]
    if codebooks is not None:
        cb = _f32(codebooks).reshape(M, ksub * dsub)
        for m in range(M):
            out.append(struct.pack("<I", ksub * dsub))
            out.append(cb[m].tobytes())
    codes = np.ascontiguousarray(codes, dtype=np.uint8).reshape(len(ids), M)
    out.append(struct.pack("<I", len(ids)))
    for i, id_ in enumerate(ids):
        out.append(struct.pack("<I", int(id_)))
        out.append(codes[i].tobytes())
    out.append(_bitmap(deleted_ids))
    return b"".join(out)


def _read_pq_params(r, M, nbits):
    got = [r.u32(n) for n in ("M", "Nbits", "Ksub", "dsub")]
    return got


def read_pq(data, dim, metric, M, nbits):
    ksub, dsub = 1 << nbits, dim // M
    r = _Reader(data)
    _read_header(r, b"PQIX", dim, metric)
    for name, have, want in zip(("M", "Nbits", "Ksub", "dsub"), _read_pq_params(r, M, nbits), (M, nbits, ksub, dsub)):
        if have != want:
            raise WireError(f"parameter {name} mismatch: index has {name}={want}, serialized data has {name}={have}")
    trained = r.u8("trained flag") == 1
    codebooks = None
    if trained:
        codebooks = np.zeros((M, ksub * dsub), np.float32)
        for m in range(M):
            sz = r.u32("codebook size")
            codebooks[m] = r.f32s(sz, "codebook data")
        codebooks = codebooks.reshape(M, ksub, dsub)
    n = r.u32("vector count")
    ids = np.zeros(n, np.uint32)
    codes = np.zeros((n, M), np.uint8)
    for i in range(n):
        ids[i] = r.u32("vector ID")
        codes[i] = np.frombuffer(r.take(M, "vector code"), dtype=np.uint8)
    deleted = _read_bitmap(r)
    return {"trained": trained, "codebooks": codebooks, "ids": ids, "codes": codes, "deleted": deleted, "consumed": r.at}


# ---- IVPQ (ivfpq_index.go:544-700 WriteTo, :702-960 ReadFrom) ---------------------------------------------------
def write_ivfpq(dim, metric, nlist, M, nbits, centroids, codebooks, lists, deleted_ids=()):
    """lists: per list an (ids, codes[n, M]) pair in list order; centroids / codebooks None when untrained."""
    ksub, dsub = 1 << nbits, dim // M
    trained = centroids is not None
    out = [_header(b"IVPQ", dim, metric), struct.pack("<I", nlist), _pq_params(M, nbits, ksub, dsub), struct.pack("<B", 1 if trained else 0)]
    if trained:
        c = _f32(centroids).reshape(nlist, dim)
        for l in range(nlist):
            out.append(struct.pack("<I", dim))
            out.append(c[l].tobytes())
        cb = _f32(codebooks).reshape(M, ksub * dsub)
        for m in range(M):
            out.append(struct.pack("<I", ksub * dsub))
            out.append(cb[m].tobytes())
    out.append(struct.pack("<I", len(lists)))
    for ids, codes in lists:
        codes = np.ascontiguousarray(codes, dtype=np.uint8).reshape(len(ids), M)
        out.append(struct.pack("<I", len(ids)))
        for i, id_ in enumerate(ids):
            out.append(struct.pack("<I", int(id_)))
            out.append(codes[i].tobytes())
    out.append(_bitmap(deleted_ids))
    return b"".join(out)


def read_ivfpq(data, dim, metric, nlist, M, nbits):
    ksub, dsub = 1 << nbits, dim // M
    r = _Reader(data)
    _read_header(r, b"IVPQ", dim, metric)
    nl = r.u32("nlist")
    got = _read_pq_params(r, M, nbits)
    for name, have, want in zip(("nlist", "M", "Nbits", "Ksub", "dsub"), [nl] + got, (nlist, M, nbits, ksub, dsub)):
        if have != want:
            raise WireError(f"parameter {name} mismatch: index has {name}={want}, serialized data has {name}={have}")
    trained = r.u8("trained flag") == 1
    centroids = codebooks = None
    if trained:
        centroids = np.zeros((nlist, dim), np.float32)
        for l in range(nlist):
            sz = r.u32("centroid size")
            centroids[l] = r.f32s(sz, "centroid data")
        codebooks = np.zeros((M, ksub * dsub), np.float32)
        for m in range(M):
            sz = r.u32("codebook size")
            codebooks[m] = r.f32s(sz, "codebook data")
        codebooks = codebooks.reshape(M, ksub, dsub)
    lists = []
    for _ in range(r.u32("list count")):
        sz = r.u32("list size")
        ids = np.zeros(sz, np.uint32)
        codes = np.zeros((sz, M), np.uint8)
        for i in range(sz):
            ids[i] = r.u32("list vector ID")
            codes[i] = np.frombuffer(r.take(M, "list vector code"), dtype=np.uint8)
        lists.append((ids, codes))
    deleted = _read_bitmap(r)
    return {"trained": trained, "centroids": centroids, "codebooks": codebooks, "lists": lists, "deleted": deleted,
            "consumed": r.at}


# ---- HNSW (hnsw_index.go:734-896 WriteTo, :918-1096 ReadFrom) ---------------------------------------------------
def level_mult(m):
    return 1.0 / math.log(float(m))                       # hnsw_index.go:206


def write_hnsw(dim, metric, m, efc, efs, max_level, entry, nodes, deleted_ids=()):
    """nodes: list of (id, level, vector[dim], [edge id arrays per layer 0..level]) in the order to be written
    (the reference iterates a Go map, i.e. any order; readers accept every order)."""
    out = [_header(b"HNSW", dim, metric), struct.pack("<IIId", m, efc, efs, level_mult(m)), struct.pack("<iI", max_level, entry),
           struct.pack("<I", len(nodes))]
    for id_, level, vec, edges in nodes:
        out.append(struct.pack("<IiI", int(id_), int(level), dim))
        out.append(_f32(vec).tobytes())
        out.append(struct.pack("<I", len(edges)))
        for e in edges:
            e = np.ascontiguousarray(e, dtype="<u4")
            out.append(struct.pack("<I", len(e)))
            out.append(e.tobytes())
    out.append(_bitmap(deleted_ids))
    return b"".join(out)


def read_hnsw(data, dim, metric, m, efc, efs):
    r = _Reader(data)
    _read_header(r, b"HNSW", dim, metric)
    gm, gefc, gefs = r.u32("M"), r.u32("efConstruction"), r.u32("efSearch")
    if gm != m:
        raise WireError(f"m parameter mismatch: index has m={m}, serialized data has m={gm}")
    if gefc != efc:
        raise WireError(f"efConstruction mismatch: index has {efc}, serialized data has {gefc}")
    if gefs != efs:
        raise WireError(f"efSearch mismatch: index has {efs}, serialized data has {gefs}")
    lm = r.f64("levelMult")
    max_level = r.i32("maxLevel")
    entry = r.u32("entryPoint")
    nodes = []
    for _ in range(r.u32("node count")):
        id_ = r.u32("node ID")
        level = r.i32("node level")
        vd = r.u32("vector dimension")
        vec = r.f32s(vd, "vector data")
        edges = []
        for _l in range(r.u32("edge layer count")):
            cnt = r.u32("edge count")
            edges.append(np.frombuffer(r.take(4 * cnt, "edge IDs"), dtype="<u4").copy())
        nodes.append((id_, level, vec, edges))
    deleted = _read_bitmap(r)
    return {"level_mult": lm, "max_level": max_level, "entry": entry, "nodes": nodes, "deleted": deleted, "consumed": r.at}


# ---- HNSWIndex.Flush (hnsw_index.go:348-430) --------------------------------------------------------------------
def hnsw_flush(nodes, entry, max_level, deleted_ids):
    """nodes: list of (id, level, vector, edges per layer) in INSERTION order.  Returns (nodes, entry, max_level).

    Phase 1 (:359-378) live nodes drop edges to deleted nodes; phase 2 (:383-412) a deleted entry point is replaced by
    a live node at maxLevel, else by a node of the highest level left, else the index is empty; phase 3 (:417-423)
    deleted nodes go.  The reference ranges over a Go map in phase 2 (random order); this restatement and the product
    both take the FIRST eligible node in insertion order, one of the outcomes the reference can produce."""
    dead = set(int(i) for i in deleted_ids)
    if not dead:
        return nodes, entry, max_level
    if entry in dead:
        found = False
        for id_, level, _v, _e in nodes:
            if id_ not in dead and level == max_level:
                entry, found = id_, True
                break
        if not found:
            best = -1
            for id_, level, _v, _e in nodes:
                if id_ not in dead and level > best:
                    best, entry = level, id_
            if best >= 0:
                max_level = best
            else:
                entry, max_level = 0, -1
    out = []
    for id_, level, vec, edges in nodes:
        if id_ in dead:
            continue
        out.append((id_, level, vec, [np.asarray([t for t in e if int(t) not in dead], dtype=np.uint32) for e in edges]))
    return out, entry, max_level
