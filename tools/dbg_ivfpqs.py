import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comet_b200 import capi
from oracle import oracle_py as O
import torch

def run(tag):
    n, d, nlist, M, nbits, metric = 7000, 32, 18, 8, 4, capi.L2
    rng = np.random.default_rng(200)
    x = rng.standard_normal((n, d)).astype(np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    o = O.IVFPQ(d, metric, nlist, M, nbits)
    o.train(x[:800].copy())
    g = capi.ShardedIVFPQIndex(d, metric, nlist, M, nbits, [0, 0, 0, 0])
    g.set_trained(o.centroids(), o.codebooks())
    o.add(ids, x.copy()); g.add(ids, x.copy())
    q = rng.standard_normal((8, d)).astype(np.float32)
    def chk(k, npb, what, nq=8, **kw):
        gi, gs, gc = g.search(q[:nq], k=k, nprobes=npb, **kw)
        torch.cuda.synchronize()
        ok = True
        for i in range(nq):
            oi, os_ = o.search(q[i], k=k, nprobes=npb, **kw)
            ok &= int(gc[i]) == len(oi) and np.array_equal(gi[i, :len(oi)], oi)
        print(tag, what, "k", k, "np", npb, "ok", ok, "counts", gc.tolist(), flush=True)
    chk(10, 4, "fresh")
    chk(200, 6, "fresh")
    chk(30, 6, "fresh")
    chk(0, 2, "k all", nq=3)
    chk(30, 6, "after k all")
    chk(50, 5, "filter", nq=3, filter_ids=np.arange(2, 7000, 4, dtype=np.uint32))
    chk(30, 6, "after filter")
    for dead in (5, 3500, 6999):
        g.remove(dead); o.remove(dead)
    chk(30, 6, "after remove")
    chk(200, 6, "after remove")
    g.flush(); o.flush()
    chk(30, 6, "after flush")

run("threads")
os.environ["COMET_B200_SHARD_THREADS"] = "0"
run("nothreads")
