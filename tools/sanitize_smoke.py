#!/usr/bin/env python
"""Small invocation of every device path, meant to run under `compute-sanitizer --tool memcheck` (and racecheck):
sizes are tiny so the 10-50x slowdown stays within a minute."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from comet_b200 import capi  # noqa: E402

rng = np.random.default_rng(0)
d = 40
x = rng.standard_normal((17000, d)).astype(np.float32)
ids = np.arange(1, len(x) + 1, dtype=np.uint32)
q = rng.standard_normal((70, d)).astype(np.float32)
for metric in (capi.L2SQ, capi.COSINE):
    f = capi.FlatIndex(d, metric)
    f.add(ids, x.copy())
    f.remove(5)
    f.search(q[:9], k=10, path=capi.PATH_EXACT)
    f.search(q, k=10, path=capi.PATH_TENSOR)
    f.search(q[:3], k=0, path=capi.PATH_EXACT)
    f.flush()
iv = capi.IVFIndex(d, 16, capi.L2)
iv.train(x[:2000].copy()); iv.add(ids[:5000], x[:5000].copy()); iv.search(q[:9], k=10, nprobes=4)
pq = capi.PQIndex(d, capi.L2, 8, 4)
pq.train(x[:2000].copy()); pq.add(ids[:5000], x[:5000].copy()); pq.search(q[:9], k=10)
ip = capi.IVFPQIndex(d, capi.L2, 8, 8, 4)
ip.train(x[:2000].copy()); ip.add(ids[:5000], x[:5000].copy()); ip.search(q[:9], k=10, nprobes=3)
# the ADC kernel's wide-row variants (M / 16 = 1 and 6: rounds of 4 / 3 rows per thread, table fast path for dsub 8),
# enough codes per (query, probe) for several rounds, the radix-select compaction and the list-ordered code copy
for (d2, M2, n2) in ((64, 16, 6000), (768, 96, 2600)):
    x2 = rng.standard_normal((n2, d2)).astype(np.float32)
    q2 = rng.standard_normal((5, d2)).astype(np.float32)
    ids2 = np.arange(1, n2 + 1, dtype=np.uint32)
    pq2 = capi.PQIndex(d2, capi.L2, M2, 8)
    pq2.train(x2[:600].copy()); pq2.add(ids2, x2.copy()); pq2.search(q2, k=20); pq2.search(q2[:1], k=300)
    ip2 = capi.IVFPQIndex(d2, capi.L2, 2, M2, 8)
    ip2.train(x2[:600].copy()); ip2.add(ids2, x2.copy()); ip2.search(q2, k=20, nprobes=2)
    ip2.remove(3); ip2.search(q2[:2], k=20, nprobes=1); ip2.flush(); ip2.search(q2[:2], k=20, nprobes=2)
    # M = 96 takes the ring form of the scan (lane-owned table banks, pre-skewed codes); the row-per-lane form again
    os.environ["COMET_B200_ADC_RING"] = "0"
    pq2.search(q2, k=20); ip2.search(q2, k=20, nprobes=2)
    os.environ.pop("COMET_B200_ADC_RING")
    ip2.search(q2, k=300, nprobes=2); ip2.search(q2[:2], k=0, nprobes=2)
# tensor path with 256-byte re-score pieces (row pitch a multiple of 64 floats) and both select launch shapes
x3 = rng.standard_normal((17000, 128)).astype(np.float32)
f3 = capi.FlatIndex(128, capi.L2)
f3.add(np.arange(1, 17001, dtype=np.uint32), x3)
q3 = rng.standard_normal((70, 128)).astype(np.float32)
f3.search(q3, k=10, path=capi.PATH_TENSOR)
os.environ["COMET_B200_SEL_SMALL"] = "1"
f3.search(q3, k=10, path=capi.PATH_TENSOR)
f3.search(q3, k=10, path=capi.PATH_TENSOR)
os.environ.pop("COMET_B200_SEL_SMALL")
# HNSW insertion on the device
hb = capi.HNSWIndex(d, capi.L2, 4, 20, 16)
hb.add(ids[:200], x[:200].copy(), np.minimum(np.floor(-np.log(1.0 - rng.random(200)) / np.log(4.0)), 5).astype(np.int32))
hb.search(q[:4], k=5)
# a hand-made HNSW graph: a ring with chords on layer 0, a few nodes on layer 1
n = 300
levels = np.zeros(n, np.int32); levels[::50] = 1
l0_off = np.arange(n + 1, dtype=np.int64) * 4
l0 = np.stack([(np.arange(n) + s) % n + 1 for s in (1, n - 1, 7, n - 7)], axis=1).astype(np.uint32).ravel()
up = np.nonzero(levels == 1)[0]
l1_off = np.zeros(n + 1, np.int64)
l1 = []
for s in range(n):
    if levels[s] == 1:
        l1.extend([int(u) + 1 for u in up if u != s][:3])
        l1_off[s + 1] = l1_off[s] + min(3, len(up) - 1)
    else:
        l1_off[s + 1] = l1_off[s]
h = capi.HNSWIndex(d, capi.L2, 4, 20, 16)
h.load_graph(ids[:n], x[:n], levels, [(l0_off, l0), (l1_off, np.asarray(l1, np.uint32))], 1, 1)
h.search(q[:9], k=5)
# round 2: row shards / list shards (one process, shards sharing the device), device filters, big-k fallbacks, wire formats,
# HNSW flush
sf = capi.ShardedFlatIndex(d, capi.L2SQ, [0, 0, 0], 6000)
sf.add(ids, x.copy()); sf.search(q[:9], k=10); sf.search(q[:2], k=0); sf.remove(7); sf.flush(); sf.search(q, k=10)
si = capi.ShardedIVFIndex(d, 16, capi.L2, [0, 0])
si.train(x[:2000].copy()); si.add(ids[:5000], x[:5000].copy()); si.search(q[:9], k=10, nprobes=4); si.rebalance(); si.search(q[:9], k=10, nprobes=4)
sp = capi.ShardedIVFPQIndex(d, capi.L2, 8, 8, 4, [0, 0])
sp.train(x[:2000].copy()); sp.add(ids[:5000], x[:5000].copy()); sp.search(q[:9], k=10, nprobes=3); sp.rebalance(); sp.search(q[:9], k=10, nprobes=3)
f = capi.FlatIndex(d, capi.L2)
f.add(ids, x.copy())
f.search(q[:9], k=10, filter_ids=ids[:300])                 # selective: candidate list
f.search(q[:9], k=10, filter_ids=ids[::3].copy())           # skip mask, exact scan
f.search(q, k=10, filter_ids=ids[::3].copy(), path=capi.PATH_TENSOR)
iv.search(q[:2], k=0, nprobes=16); pq.search(q[:2], k=0); ip.search(q[:2], k=0, nprobes=8)
for kind, ix, mk in (("flat", f, lambda: capi.FlatIndex(d, capi.L2)), ("ivf", iv, lambda: capi.IVFIndex(d, 16, capi.L2)),
                     ("pq", pq, lambda: capi.PQIndex(d, capi.L2, 8, 4)), ("ivfpq", ip, lambda: capi.IVFPQIndex(d, capi.L2, 8, 8, 4))):
    blob = capi.save_bytes(kind, ix.h)
    other = mk()
    capi.load_bytes(kind, other.h, blob)
    other.search(q[:3], k=5)
hb.remove(int(ids[0])); hb.remove(int(ids[3])); hb.flush(); hb.search(q[:4], k=5)
blob = capi.save_bytes("hnsw", hb.h)
hb2 = capi.HNSWIndex(d, capi.L2, 4, 20, 16)
capi.load_bytes("hnsw", hb2.h, blob); hb2.search(q[:4], k=5)
print("sanitize smoke done")
