#!/usr/bin/env python
"""IVF / IVFPQ list shards over the GPUs of one box (cm_ivf_sharded_*, cm_ivfpq_sharded_*): time per batch against the
single-GPU index on the same data, bit-equality of the answers, and the per-shard scanned vectors (the balance of the
l mod W assignment and of the greedy-by-length one after rebalance)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from comet_b200 import capi  # noqa: E402


def timed(fn, reps=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2_000_000)
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--nlist", type=int, default=1024)
    args = ap.parse_args()
    W = args.gpus or torch.cuda.device_count()
    d, nq, k, nprobe, M = 768, 512, 100, 32, 96
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(3)
    W3 = torch.randn((32, d), generator=g, device=dev)

    def rows(m):
        z = torch.randn((m, 32), generator=g, device=dev)
        return (z @ W3 + 0.05 * torch.randn((m, d), generator=g, device=dev)).cpu().numpy()

    train = rows(args.nlist * 16)
    q = rows(nq)
    out = {"gpus": W, "n": args.n, "dim": d, "nlist": args.nlist, "nprobes": nprobe, "k": k, "nq": nq}
    one_ivf = capi.IVFIndex(d, args.nlist, capi.L2)
    one_ivf.train(train.copy())
    sh_ivf = capi.ShardedIVFIndex(d, args.nlist, capi.L2, list(range(W)))
    sh_ivf.set_centroids(one_ivf.centroids())
    one_pq = capi.IVFPQIndex(d, capi.L2, args.nlist, M, 8)
    one_pq.train(train.copy())
    sh_pq = capi.ShardedIVFPQIndex(d, capi.L2, args.nlist, M, 8, list(range(W)))
    sh_pq.set_trained(*one_pq.trained_state())
    for s0 in range(0, args.n, 250_000):
        m = min(250_000, args.n - s0)
        x = rows(m)
        ids = np.arange(s0 + 1, s0 + m + 1, dtype=np.uint32)
        for ix in (one_ivf, sh_ivf, one_pq, sh_pq):
            ix.add(ids, x.copy(), writeback=False)
    for name, one, sh in (("ivf", one_ivf, sh_ivf), ("ivfpq", one_pq, sh_pq)):
        a = one.search(q, k=k, nprobes=nprobe)
        b = sh.search(q, k=k, nprobes=nprobe)
        same = bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32)) and np.array_equal(a[2], b[2]))
        t_one = timed(lambda: one.search(q, k=k, nprobes=nprobe))
        t_sh = timed(lambda: sh.search(q, k=k, nprobes=nprobe))
        scanned = sh.last_scanned().tolist()
        sizes = [sh.shard_size(r) for r in range(W)]
        sh.rebalance()
        b2 = sh.search(q, k=k, nprobes=nprobe)
        same2 = bool(np.array_equal(a[0], b2[0]) and np.array_equal(a[1].view(np.uint32), b2[1].view(np.uint32)))
        t_sh2 = timed(lambda: sh.search(q, k=k, nprobes=nprobe))
        scanned2 = sh.last_scanned().tolist()
        sizes2 = [sh.shard_size(r) for r in range(W)]
        out[name] = {"single_gpu_ms": t_one * 1e3, "sharded_ms_mod_assignment": t_sh * 1e3, "sharded_ms_greedy_assignment": t_sh2 * 1e3,
                     "bit_identical_to_single_index": same and same2,
                     "mod_assignment": {"vectors_per_shard": sizes, "scanned_per_shard": scanned,
                                        "scan_imbalance_max_over_mean": max(scanned) / (sum(scanned) / W)},
                     "greedy_assignment": {"vectors_per_shard": sizes2, "scanned_per_shard": scanned2,
                                           "scan_imbalance_max_over_mean": max(scanned2) / (sum(scanned2) / W)}}
        print(name, json.dumps(out[name]), file=sys.stderr, flush=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
