import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comet_b200 import capi
L = capi.lib()
rng = np.random.default_rng(1)
n, d, nq = 200_000, 768, 256
x = rng.standard_normal((n, d), dtype=np.float32); q = rng.standard_normal((nq, d), dtype=np.float32)
ids = np.arange(1, n + 1, dtype=np.uint32)
for which in sys.argv[1:]:
    if which == "ivf":
        ix = capi.IVFIndex(d, 512, capi.L2); ix.train(x[:4096].copy()); ix.add(ids, x.copy(), writeback=False)
        f = lambda: ix.search(q, k=100, nprobes=16)
    elif which == "pq":
        ix = capi.PQIndex(d, capi.L2, 96, 8); ix.train(x[:4096].copy()); ix.add(ids, x.copy(), writeback=False)
        f = lambda: ix.search(q[:64], k=100)
    else:
        ix = capi.IVFPQIndex(d, capi.L2, 512, 96, 8); ix.train(x[:5120].copy()); ix.add(ids, x.copy(), writeback=False)
        f = lambda: ix.search(q, k=100, nprobes=16)
    f(); f()
    L.cm_profile_reset(); L.cm_profile_enable(1)
    k0 = L.cm_kernel_launches()
    t0 = time.perf_counter(); f(); wall = (time.perf_counter() - t0) * 1e3
    L.cm_profile_enable(0)
    parts = {name: capi.profile_get(c) for name, c in [("flat_scan", 0), ("select", 3), ("ivf_scan", 4), ("pq_scan", 5)]}
    sc = L.cm_ivf_last_scanned(ix.h) if which == "ivf" else (L.cm_ivfpq_last_scanned(ix.h) if which == "ivfpq" else 64 * n)
    print(which, "scanned", sc, "wall ms %.2f" % wall, "launches", L.cm_kernel_launches() - k0, {k: (round(v[0], 3), v[1]) for k, v in parts.items()}, flush=True)
    del ix
