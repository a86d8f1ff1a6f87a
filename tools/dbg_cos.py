import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comet_b200 import capi
def run(metric, n=33000, d=128, nq=260, k=10):
    rng = np.random.default_rng(11 + metric)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    g = capi.FlatIndex(d, metric)
    g.add(ids, x.copy())
    q = rng.standard_normal((nq, d)).astype(np.float32)
    for rep in range(3):
        i, s, c = g.search(q, k=k, path=capi.PATH_TENSOR)
        st = g.last_stats()
        print("metric", metric, "rep", rep, "fallback", st["fallback_queries"], "cand", st["candidates"], "cnt min", c.min(), flush=True)
os.environ["COMET_B200_DBG_STAGED"] = "1"
for m in (1, 2, 0):
    run(m)
