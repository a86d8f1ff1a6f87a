"""The oracle's restatement of the reference's wire formats (oracle/wire_py.py), pinned by hand-built streams:
every stream below is written out field by field from the reference's format comments (flat_index.go:348-359,
ivf_index.go:440-461, pq_index.go:480-502, ivfpq_index.go:507-537, hnsw_index.go:701-727) with struct.pack, NOT
with the writer under test.  Also the roaring portable format (RoaringFormatSpec) and the library's host-side
roaring decoder (no GPU involved)."""
import struct

import numpy as np
import pytest

from comet_b200 import capi
from oracle import wire_py as W

EMPTY = struct.pack("<I", 8) + bytes([0x3A, 0x30, 0, 0, 0, 0, 0, 0])     # length + empty roaring bitmap


def f32(*v):
    return struct.pack("<%df" % len(v), *v)


def test_flat_stream_field_by_field():
    stream = (b"FLAT" + struct.pack("<I", 1) + struct.pack("<I", 3) + struct.pack("<I", 2) + b"l2"
              + struct.pack("<I", 2)
              + struct.pack("<II", 7, 3) + f32(1.0, 2.0, 3.0)
              + struct.pack("<II", 9, 3) + f32(-0.5, 0.25, 4.0)
              + EMPTY)
    ids = np.array([7, 9], np.uint32)
    rows = np.array([[1, 2, 3], [-0.5, 0.25, 4]], np.float32)
    assert W.write_flat(3, 0, ids, rows) == stream
    got = W.read_flat(stream, 3, 0)
    assert got["ids"].tolist() == [7, 9] and np.array_equal(got["rows"], rows) and got["deleted"] == []
    assert got["consumed"] == len(stream) == 4 + 4 + 4 + 4 + 2 + 4 + 2 * (8 + 12) + 12


def test_ivf_stream_field_by_field():
    stream = (b"IVFX" + struct.pack("<II", 1, 2) + struct.pack("<I", 6) + b"cosine"
              + struct.pack("<I", 2) + b"\x01"
              + struct.pack("<I", 2) + f32(1.0, 0.0) + struct.pack("<I", 2) + f32(0.0, 1.0)
              + struct.pack("<I", 2)
              + struct.pack("<I", 1) + struct.pack("<I", 11) + f32(0.6, 0.8)
              + struct.pack("<I", 2) + struct.pack("<I", 12) + f32(0.0, 1.0) + struct.pack("<I", 13) + f32(-1.0, 0.0)
              + EMPTY)
    cent = np.array([[1, 0], [0, 1]], np.float32)
    lists = [(np.array([11], np.uint32), np.array([[0.6, 0.8]], np.float32)),
             (np.array([12, 13], np.uint32), np.array([[0, 1], [-1, 0]], np.float32))]
    assert W.write_ivf(2, 2, 2, cent, lists) == stream
    got = W.read_ivf(stream, 2, 2, 2)
    assert got["trained"] and np.array_equal(got["centroids"], cent)
    assert [l[0].tolist() for l in got["lists"]] == [[11], [12, 13]]
    assert np.array_equal(got["lists"][1][1], lists[1][1]) and got["consumed"] == len(stream)
    # untrained: no centroid block at all
    un = b"IVFX" + struct.pack("<II", 1, 2) + struct.pack("<I", 6) + b"cosine" + struct.pack("<I", 2) + b"\x00" + struct.pack("<I", 2) + struct.pack("<II", 0, 0) + EMPTY
    assert W.write_ivf(2, 2, 2, None, [(np.zeros(0, np.uint32), np.zeros((0, 2), np.float32))] * 2) == un
    assert not W.read_ivf(un, 2, 2, 2)["trained"]


def test_pq_stream_field_by_field():
    # dim 4, M 2, Nbits 1 -> Ksub 2, dsub 2
    cb = np.arange(8, dtype=np.float32).reshape(2, 2, 2)
    stream = (b"PQIX" + struct.pack("<II", 1, 4) + struct.pack("<I", 10) + b"l2_squared"
              + struct.pack("<IIII", 2, 1, 2, 2) + b"\x01"
              + struct.pack("<I", 4) + f32(0, 1, 2, 3) + struct.pack("<I", 4) + f32(4, 5, 6, 7)
              + struct.pack("<I", 3)
              + struct.pack("<I", 5) + bytes([0, 1]) + struct.pack("<I", 6) + bytes([1, 1]) + struct.pack("<I", 8) + bytes([1, 0])
              + EMPTY)
    ids = np.array([5, 6, 8], np.uint32)
    codes = np.array([[0, 1], [1, 1], [1, 0]], np.uint8)
    assert W.write_pq(4, 1, 2, 1, cb, ids, codes) == stream
    got = W.read_pq(stream, 4, 1, 2, 1)
    assert np.array_equal(got["codebooks"], cb) and got["ids"].tolist() == [5, 6, 8] and np.array_equal(got["codes"], codes)
    with pytest.raises(W.WireError, match="parameter M mismatch: index has M=4, serialized data has M=2"):
        W.read_pq(stream, 4, 1, 4, 1)


def test_ivfpq_stream_field_by_field():
    cb = np.arange(8, dtype=np.float32).reshape(2, 2, 2)
    cent = np.array([[1, 1, 1, 1], [2, 2, 2, 2], [3, 3, 3, 3]], np.float32)
    stream = (b"IVPQ" + struct.pack("<II", 1, 4) + struct.pack("<I", 2) + b"l2"
              + struct.pack("<I", 3) + struct.pack("<IIII", 2, 1, 2, 2) + b"\x01"
              + b"".join(struct.pack("<I", 4) + f32(*cent[l]) for l in range(3))
              + struct.pack("<I", 4) + f32(0, 1, 2, 3) + struct.pack("<I", 4) + f32(4, 5, 6, 7)
              + struct.pack("<I", 3)
              + struct.pack("<I", 0)
              + struct.pack("<I", 2) + struct.pack("<I", 21) + bytes([1, 0]) + struct.pack("<I", 22) + bytes([0, 0])
              + struct.pack("<I", 1) + struct.pack("<I", 23) + bytes([1, 1])
              + EMPTY)
    lists = [(np.zeros(0, np.uint32), np.zeros((0, 2), np.uint8)),
             (np.array([21, 22], np.uint32), np.array([[1, 0], [0, 0]], np.uint8)),
             (np.array([23], np.uint32), np.array([[1, 1]], np.uint8))]
    assert W.write_ivfpq(4, 0, 3, 2, 1, cent, cb, lists) == stream
    got = W.read_ivfpq(stream, 4, 0, 3, 2, 1)
    assert np.array_equal(got["centroids"], cent) and np.array_equal(got["codebooks"], cb)
    assert [l[0].tolist() for l in got["lists"]] == [[], [21, 22], [23]] and got["consumed"] == len(stream)


def test_hnsw_stream_field_by_field():
    import math
    stream = (b"HNSW" + struct.pack("<II", 1, 2) + struct.pack("<I", 2) + b"l2"
              + struct.pack("<III", 4, 20, 10) + struct.pack("<d", 1.0 / math.log(4.0))
              + struct.pack("<i", 1) + struct.pack("<I", 100)
              + struct.pack("<I", 2)
              + struct.pack("<IiI", 100, 1, 2) + f32(0.5, 1.5) + struct.pack("<I", 2)
              + struct.pack("<I", 1) + struct.pack("<I", 101) + struct.pack("<I", 0)
              + struct.pack("<IiI", 101, 0, 2) + f32(2.5, -1.0) + struct.pack("<I", 1)
              + struct.pack("<I", 1) + struct.pack("<I", 100)
              + EMPTY)
    nodes = [(100, 1, np.array([0.5, 1.5], np.float32), [np.array([101], np.uint32), np.zeros(0, np.uint32)]),
             (101, 0, np.array([2.5, -1.0], np.float32), [np.array([100], np.uint32)])]
    assert W.write_hnsw(2, 0, 4, 20, 10, 1, 100, nodes) == stream
    got = W.read_hnsw(stream, 2, 0, 4, 20, 10)
    assert got["max_level"] == 1 and got["entry"] == 100 and got["level_mult"] == 1.0 / math.log(4.0)
    assert [(n[0], n[1]) for n in got["nodes"]] == [(100, 1), (101, 0)]
    assert got["nodes"][0][3][0].tolist() == [101] and got["nodes"][0][3][1].tolist() == [] and got["consumed"] == len(stream)
    with pytest.raises(W.WireError, match="efSearch mismatch: index has 11, serialized data has 10"):
        W.read_hnsw(stream, 2, 0, 4, 20, 11)


def test_header_errors_carry_the_reference_messages():
    good = W.write_flat(3, 0, np.array([1], np.uint32), np.ones((1, 3), np.float32))
    with pytest.raises(W.WireError, match="invalid magic number: expected 'FLAT', got 'FLAX'"):
        W.read_flat(b"FLAX" + good[4:], 3, 0)
    with pytest.raises(W.WireError, match="unsupported version: 2"):
        W.read_flat(good[:4] + struct.pack("<I", 2) + good[8:], 3, 0)
    with pytest.raises(W.WireError, match="dimension mismatch: index has dim=4, serialized data has dim=3"):
        W.read_flat(good, 4, 0)
    with pytest.raises(W.WireError, match="distance kind mismatch: index uses 'cosine', serialized data uses 'l2'"):
        W.read_flat(good, 3, 2)
    with pytest.raises(W.WireError, match="unexpected EOF"):
        W.read_flat(good[:-3], 3, 0)


# ---- roaring ---------------------------------------------------------------------------------------------------
def test_roaring_hand_built_blobs():
    # array container {1, 2, 3} under key 0: cookie 12346, 1 container, (key 0, card-1 2), offset 16, three u16
    blob = struct.pack("<II", 12346, 1) + struct.pack("<HH", 0, 2) + struct.pack("<I", 16) + struct.pack("<HHH", 1, 2, 3)
    assert W.roaring_decode(blob) == [1, 2, 3] and W.roaring_encode([3, 1, 2]) == blob
    # two keys: {5} and {65536 + 7}
    blob2 = (struct.pack("<II", 12346, 2) + struct.pack("<HHHH", 0, 0, 1, 0) + struct.pack("<II", 24, 26)
             + struct.pack("<H", 5) + struct.pack("<H", 7))
    assert W.roaring_decode(blob2) == [5, 65543] and W.roaring_encode([65543, 5]) == blob2
    # run container: cookie 12347 | (n-1) << 16, run flag byte, header, (no offsets below 4 containers), 1 run [10, 14]
    blob3 = struct.pack("<I", 12347) + b"\x01" + struct.pack("<HH", 0, 4) + struct.pack("<H", 1) + struct.pack("<HH", 10, 4)
    assert W.roaring_decode(blob3) == [10, 11, 12, 13, 14] and W.roaring_encode(range(10, 15), use_runs=True) == blob3
    assert W.roaring_decode(W.EMPTY_ROARING) == [] and W.roaring_decode(b"") == []


@pytest.mark.parametrize("use_runs", [False, True])
def test_roaring_round_trip_and_the_library_decoder(use_runs):
    rng = np.random.default_rng(3)
    sets = [
        [],
        [0],
        [4294967295, 0, 65535, 65536],
        rng.choice(200000, size=3000, replace=False).tolist(),            # several array containers
        (np.arange(70000) * 1).tolist(),                                  # a full 65536 bitmap container + an array
        (rng.choice(65536, size=5000, replace=False) + 3 * 65536).tolist(),    # one bitmap container (card > 4096)
        list(range(100, 9000)) + list(range(70000, 70010)) + [5 * 65536 + i for i in range(0, 3000, 3)]
        + [9 * 65536 + 1, 12 * 65536 + 2],                                # 5 containers: run form carries offsets
    ]
    for ids in sets:
        blob = W.roaring_encode(ids, use_runs=use_runs) if ids else W.EMPTY_ROARING
        want = sorted(set(ids))
        assert W.roaring_decode(blob) == want
        assert capi.decode_roaring(blob).tolist() == want                 # the product's host-side decoder


def test_library_decoder_rejects_garbage():
    with pytest.raises(capi.CometError):
        capi.decode_roaring(struct.pack("<II", 999, 1))
    with pytest.raises(capi.CometError):
        capi.decode_roaring(struct.pack("<II", 12346, 1) + struct.pack("<HH", 0, 2))      # truncated


# ---- HNSWIndex.Flush restatement ---------------------------------------------------------------------------------
def _nodes():
    v = np.zeros(2, np.float32)
    u = lambda *a: np.array(a, np.uint32)     # noqa: E731
    return [(1, 2, v, [u(2, 3), u(2), u()]),
            (2, 1, v, [u(1, 3, 4), u(1)]),
            (3, 0, v, [u(1, 2, 4)]),
            (4, 1, v, [u(2, 3), u()])]


def test_hnsw_flush_restatement():
    # nothing deleted: untouched
    assert W.hnsw_flush(_nodes(), 1, 2, [])[1:] == (1, 2)
    # a non-entry node goes: its edges disappear everywhere, entry stays
    nodes, entry, ml = W.hnsw_flush(_nodes(), 1, 2, [3])
    assert [n[0] for n in nodes] == [1, 2, 4] and entry == 1 and ml == 2
    assert nodes[0][3][0].tolist() == [2] and nodes[1][3][0].tolist() == [1, 4] and nodes[2][3][0].tolist() == [2]
    # the entry point goes, nobody else at maxLevel: the highest level left wins, first in insertion order (2 before 4)
    nodes, entry, ml = W.hnsw_flush(_nodes(), 1, 2, [1])
    assert entry == 2 and ml == 1 and [n[0] for n in nodes] == [2, 3, 4]
    assert nodes[0][3][0].tolist() == [3, 4] and nodes[0][3][1].tolist() == []
    # everything goes
    assert W.hnsw_flush(_nodes(), 1, 2, [1, 2, 3, 4]) == ([], 0, -1)
    # entry point deleted but another node sits at maxLevel
    n2 = _nodes()
    n2[3] = (4, 2, n2[3][2], [n2[3][3][0], n2[3][3][1], np.zeros(0, np.uint32)])
    assert W.hnsw_flush(n2, 1, 2, [1])[1:] == (4, 2)
