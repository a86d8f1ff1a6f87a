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
print("sanitize smoke done")
