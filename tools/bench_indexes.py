#!/usr/bin/env python
"""Throughput of the IVF / PQ / IVFPQ / HNSW device search paths at moderate sizes (the BASELINE.json
configs[2..3] are parity-test cases; this script characterises their kernels for DESIGN.md / profiles/).
Indexes are trained, filled and -- for HNSW -- built through the C ABI on the GPU."""
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


def full_size(args):
    """BASELINE.json configs[2] and configs[3] at their full sizes.  Rows are generated on the GPU with torch
    (plumbing: a host numpy generator would dominate the box time) and handed to the C ABI's host entry points
    in slabs, exactly like a caller that streams its corpus in."""
    import torch
    out = {}
    d, nq = 768, 512
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(20261017)
    only = set(args.only.split(","))

    def blobs(centers, m):
        idx = torch.randint(0, centers.shape[0], (m,), generator=g, device=dev)
        return (centers[idx] + torch.randn((m, d), generator=g, device=dev)).cpu().numpy()

    if "c3" in only:
        n, nlist, nprobe, M = args.c3_n, args.c3_nlist, 32, 96
        # fewer blobs than lists: every blob is represented among the 65,536 training rows, so no blob's rows fall
        # into a far-away "hub" list at Add time (with 8,192 blobs a query scanned 18 % of the corpus)
        centers = torch.randn((args.c3_blobs, d), generator=g, device=dev) * 2.0
        if args.c3_data == "manifold":
            # rows on a 32-dimensional linear manifold + small isotropic noise: one connected cloud, so the
            # reference's deterministic k-means yields lists of comparable length (no blob / hub structure)
            W3 = torch.randn((32, d), generator=g, device=dev)

            def blobs(_c, m):   # noqa: F811
                z = torch.randn((m, 32), generator=g, device=dev)
                return (z @ W3 + 0.05 * torch.randn((m, d), generator=g, device=dev)).cpu().numpy()
        ix = capi.IVFPQIndex(d, capi.L2, nlist, M, 8)
        t0 = time.perf_counter(); ix.train(blobs(centers, nlist * 16)); t_train = time.perf_counter() - t0
        t_add = 0.0
        slab = 500_000
        for s0 in range(0, n, slab):
            m = min(slab, n - s0)
            x = blobs(centers, m)
            t0 = time.perf_counter()
            ix.add(np.arange(s0 + 1, s0 + m + 1, dtype=np.uint32), x, writeback=False)
            t_add += time.perf_counter() - t0
            del x
        q = blobs(centers, nq)
        L = capi.lib()
        L.cm_profile_reset(); L.cm_profile_enable(1)
        dt = timed(lambda: ix.search(q, k=100, nprobes=nprobe), reps=5)
        L.cm_profile_enable(0)
        scan_ms, scan_n = capi.profile_get(capi.PROF_PQ_SCAN)
        scanned = L.cm_ivfpq_last_scanned(ix.h) / nq
        per = scan_ms / max(scan_n, 1) * 1e-3
        out["c3_ivfpq"] = {"data": args.c3_data, "n": n, "dim": d, "nlist": nlist, "nprobes": nprobe, "M": M, "nbits": 8, "k": 100, "nq": nq,
                           "train_s": t_train, "add_s": t_add, "qps_host_api": nq / dt, "ms_per_batch": dt * 1e3,
                           "scanned_per_query": scanned, "adc_kernel_ms": per * 1e3,
                           "adc_lookups_per_s": nq * scanned * M / per, "adc_code_GBps": nq * scanned * (M + 4) / per / 1e9,
                           "lut_builds_per_s": nq * nprobe / per}
        # single query latency (the reference's Execute() shape)
        dt1 = timed(lambda: ix.search(q[:1], k=100, nprobes=nprobe), reps=20)
        out["c3_ivfpq"]["ms_single_query"] = dt1 * 1e3
        del ix
        print(json.dumps(out), file=sys.stderr, flush=True)
    if "c4" in only:
        hn = args.c4_n
        ids = np.arange(1, hn + 1, dtype=np.uint32)
        Wk = torch.randn((24, d), generator=g, device=dev)
        xk = (torch.randn((hn, 24), generator=g, device=dev) @ Wk).cpu().numpy()
        qk = (torch.randn((nq, 24), generator=g, device=dev) @ Wk).cpu().numpy()
        flat = capi.FlatIndex(d, capi.L2)
        flat.add(ids, xk.copy())
        nbr = np.zeros((hn, 32), np.uint32)
        t0 = time.perf_counter()
        for s0 in range(0, hn, 16384):
            gi, _, _ = flat.search(xk[s0:s0 + 16384], k=33)
            nbr[s0:s0 + 16384] = gi[:, 1:33]
        t_knn = time.perf_counter() - t0
        ti, _, _ = flat.search(qk, k=10)
        del flat
        levels = np.zeros(hn, np.int32)
        off = np.arange(hn + 1, dtype=np.int64) * 32
        gidx = capi.HNSWIndex(d, capi.L2, 16, 100, 128)
        gidx.load_graph(ids, xk, levels, [(off, nbr.ravel())], 1, 0)
        res = {}

        def run_knn():
            res["r"] = gidx.search(qk, k=10, ef_search=128, with_work=True)
        dt = timed(run_knn)
        work = res["r"][3]
        evals = float(work[:, 0].mean())
        rec = float(np.mean([len(set(res["r"][0][i, :10].tolist()) & set(ti[i].tolist())) / 10 for i in range(nq)]))
        # the traversal is a latency-bound chain per query: throughput comes from queries in flight
        sweep = {}
        for nq2 in [int(v) for v in args.c4_nq.split(",") if v]:
            qq = (torch.randn((nq2, 24), generator=g, device=dev) @ Wk).cpu().numpy()
            dt2 = timed(lambda: gidx.search(qq, k=10, ef_search=128), reps=3)
            sweep[str(nq2)] = {"ms_per_batch": dt2 * 1e3, "qps_host_api": nq2 / dt2}
        out["c4_hnsw_knn_graph"] = {"queries_in_flight_sweep": sweep, "n": hn, "dim": d, "degree": 32, "ef": 128, "k": 10, "nq": nq, "knn_graph_build_s": t_knn,
                                    "qps_host_api": nq / dt, "ms_per_batch": dt * 1e3, "dist_evals_per_query": evals,
                                    "expansions_per_query": float(work[:, 1].mean()),
                                    "algorithmic_GBps": nq * evals * (d * 4 + 4) / dt / 1e9, "recall_at_10_vs_flat": rec}
    print(json.dumps(out, indent=1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=500_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--nq", type=int, default=512)
    ap.add_argument("--hnsw-n", type=int, default=20_000)
    ap.add_argument("--only", default="", help="comma list of ivf,pq,ivfpq,hnsw,hnswknn,c3,c4")
    ap.add_argument("--c3-n", type=int, default=10_000_000, help="rows of the BASELINE configs[2] run (IVFPQ)")
    ap.add_argument("--c4-nq", default="2048,8192", help="extra batch sizes for the HNSW run")
    ap.add_argument("--c3-blobs", type=int, default=1024)
    ap.add_argument("--c3-nlist", type=int, default=4096)
    ap.add_argument("--c3-data", default="manifold", choices=["manifold", "blobs"])
    ap.add_argument("--c4-n", type=int, default=1_000_000, help="rows of the BASELINE configs[3] run (HNSW)")
    args = ap.parse_args()
    if args.only in ("c3", "c4", "c3,c4"):
        return full_size(args)
    only = set(args.only.split(",")) if args.only else {"ivf", "pq", "ivfpq", "hnsw"}
    rng = np.random.default_rng(1)
    n, d, nq = args.n, args.dim, args.nq
    # clustered corpus (2048 Gaussian blobs): i.i.d. N(0,1) rows have no cluster structure in 768-d, k-means
    # lists come out wildly uneven and every query probes the few giant lists
    centers = rng.standard_normal((2048, d), dtype=np.float32) * 2.0
    x = centers[rng.integers(0, 2048, n)] + rng.standard_normal((n, d), dtype=np.float32)
    q = centers[rng.integers(0, 2048, nq)] + rng.standard_normal((nq, d), dtype=np.float32)
    ids = np.arange(1, n + 1, dtype=np.uint32)
    out = {}

    nlist, nprobe, M = 1024, 32, 96
    if "ivf" in only:
        ivf = capi.IVFIndex(d, nlist, capi.L2)
        t0 = time.perf_counter(); ivf.train(x[: nlist * 16].copy()); t_train = time.perf_counter() - t0
        t0 = time.perf_counter(); ivf.add(ids, x.copy(), writeback=False); t_add = time.perf_counter() - t0
        r_lm = ivf.search(q, k=100, nprobes=nprobe)
        dt = timed(lambda: ivf.search(q, k=100, nprobes=nprobe))
        scanned = capi.lib().cm_ivf_last_scanned(ivf.h) / nq       # measured: vectors of probed lists per query
        L = capi.lib()
        L.cm_profile_reset(); L.cm_profile_enable(1)
        ivf.search(q, k=100, nprobes=nprobe)
        L.cm_profile_enable(0)
        phases = {"list_scan_ms": capi.profile_get(capi.PROF_IVF_SCAN)[0], "coarse_scan_ms": capi.profile_get(capi.PROF_FLAT_SCAN)[0],
                  "merges_ms": capi.profile_get(capi.PROF_SELECT)[0]}
        os.environ["COMET_B200_IVF_LIST_MAJOR"] = "0"              # the query-major scan on the same index
        r_qm = ivf.search(q, k=100, nprobes=nprobe)
        dt_qm = timed(lambda: ivf.search(q, k=100, nprobes=nprobe), reps=3)
        os.environ.pop("COMET_B200_IVF_LIST_MAJOR")
        same = all(np.array_equal(a, b) for a, b in zip(r_lm, r_qm))
        out["ivf"] = {"n": n, "dim": d, "nlist": nlist, "nprobes": nprobe, "k": 100, "train_s": t_train, "add_s": t_add,
                      "qps_host_api": nq / dt, "ms_per_batch": dt * 1e3, "scanned_per_query": scanned,
                      "algorithmic_GBps": nq * (scanned * d * 4 + nlist * d * 4 / 8) / dt / 1e9,
                      "phases_of_one_search": phases,
                      "query_major_scan": {"ms_per_batch": dt_qm * 1e3, "qps_host_api": nq / dt_qm, "same_results": bool(same)}}
        del ivf
    if "pq" in only:
        pq = capi.PQIndex(d, capi.L2, M, 8)
        t0 = time.perf_counter(); pq.train(x[:20000].copy()); t_train = time.perf_counter() - t0
        t0 = time.perf_counter(); pq.add(ids, x.copy(), writeback=False); t_add = time.perf_counter() - t0
        r_ring = pq.search(q[:128], k=100)
        dt = timed(lambda: pq.search(q[:128], k=100), reps=3)
        os.environ["COMET_B200_ADC_RING"] = "0"          # the row-per-lane form on the same index
        r_rows = pq.search(q[:128], k=100)
        dt_rows = timed(lambda: pq.search(q[:128], k=100), reps=3)
        os.environ.pop("COMET_B200_ADC_RING")
        same = all(np.array_equal(a, b) for a, b in zip(r_ring, r_rows))
        out["pq"] = {"n": n, "dim": d, "M": M, "nbits": 8, "k": 100, "train_s": t_train, "add_s": t_add, "qps_host_api": 128 / dt,
                     "ms_per_batch_128q": dt * 1e3, "lookups_per_s": 128 * n * M / dt, "code_GBps": 128 * n * M / dt / 1e9,
                     "row_per_lane_form": {"ms_per_batch_128q": dt_rows * 1e3, "lookups_per_s": 128 * n * M / dt_rows,
                                           "same_results": bool(same)}}
        del pq
    if "ivfpq" in only:
        ivfpq = capi.IVFPQIndex(d, capi.L2, nlist, M, 8)
        t0 = time.perf_counter(); ivfpq.train(x[: nlist * 16].copy()); t_train = time.perf_counter() - t0
        t0 = time.perf_counter(); ivfpq.add(ids, x.copy(), writeback=False); t_add = time.perf_counter() - t0
        dt = timed(lambda: ivfpq.search(q, k=100, nprobes=nprobe))
        scanned = capi.lib().cm_ivfpq_last_scanned(ivfpq.h) / nq
        out["ivfpq"] = {"n": n, "dim": d, "nlist": nlist, "nprobes": nprobe, "M": M, "nbits": 8, "k": 100, "train_s": t_train,
                        "add_s": t_add, "qps_host_api": nq / dt, "ms_per_batch": dt * 1e3, "scanned_per_query": scanned,
                        "lookups_per_s": nq * scanned * M / dt, "lut_builds_per_s": nq * nprobe / dt}
        del ivfpq
    if "hnswknn" in only:
        # A HEALTHY layer-0 graph (exact 32-nearest-neighbour graph built with the flat index on the GPU) to show
        # what the traversal kernel does when the search really explores: graphs built by the reference's own
        # insertNode collapse to a 2M+1 clique around the entry point (see "hnsw" below and DESIGN.md).
        hn = min(n, 200_000)
        # rows on a 24-dimensional manifold (one connected cloud; the blob corpus above gives one component per blob)
        Wk = rng.standard_normal((24, d), dtype=np.float32)
        xk = (rng.standard_normal((hn, 24), dtype=np.float32) @ Wk).astype(np.float32)
        qk = (rng.standard_normal((nq, 24), dtype=np.float32) @ Wk).astype(np.float32)
        flat = capi.FlatIndex(d, capi.L2)
        flat.add(ids[:hn], xk.copy())
        nbr = np.zeros((hn, 32), np.uint32)
        for s0 in range(0, hn, 4096):
            gi, _, _ = flat.search(xk[s0:s0 + 4096], k=33)
            nbr[s0:s0 + 4096] = gi[:, 1:33]
        del flat
        levels = np.zeros(hn, np.int32)
        off = np.arange(hn + 1, dtype=np.int64) * 32
        g = capi.HNSWIndex(d, capi.L2, 16, 100, 128)
        g.load_graph(ids[:hn], xk, levels, [(off, nbr.ravel())], 1, 0)
        res = {}

        def run_knn():
            res["r"] = g.search(qk, k=10, ef_search=128, with_work=True)
        dt = timed(run_knn)
        work = res["r"][3]
        evals = float(work[:, 0].mean())
        flat2 = capi.FlatIndex(d, capi.L2)
        flat2.add(ids[:hn], xk.copy())
        ti, _, _ = flat2.search(qk, k=10)
        rec = float(np.mean([len(set(res["r"][0][i, :10].tolist()) & set(ti[i].tolist())) / 10 for i in range(nq)]))
        out["hnsw_knn_graph"] = {"n": hn, "dim": d, "degree": 32, "ef": 128, "k": 10, "qps_host_api": nq / dt, "ms_per_batch": dt * 1e3,
                                 "dist_evals_per_query": evals, "expansions_per_query": float(work[:, 1].mean()),
                                 "algorithmic_GBps": nq * evals * (d * 4 + 4) / dt / 1e9, "recall_at_10_vs_flat": rec}
        del g, flat2
    if "hnsw" in only:
        # HNSW: graph from the host builder, search on the device.  Rows on a 24-dimensional manifold: with i.i.d.
        # 768-d Gaussian rows all pairwise distances coincide and the reference's build (entry point never promoted,
        # plain "M nearest" selection) leaves only a few dozen nodes reachable from the entry point.
        hn = args.hnsw_n
        z = rng.standard_normal((hn, 24), dtype=np.float32)
        W = rng.standard_normal((24, d), dtype=np.float32)
        xh = (z @ W + 0.05 * rng.standard_normal((hn, d), dtype=np.float32)).astype(np.float32)
        qh = (rng.standard_normal((nq, 24), dtype=np.float32) @ W).astype(np.float32)
        # the reference's level draw, floor(-ln(U) / ln(M)) (hnsw_index.go:474-484), with numpy's generator
        lv = np.minimum(np.floor(-np.log(1.0 - rng.random(hn)) / np.log(16.0)), 16).astype(np.int32)
        lv[0] = 0
        g = capi.HNSWIndex(d, capi.L2, 16, 100, 128)
        t0 = time.perf_counter(); g.add(ids[:hn], xh.copy(), lv, writeback=False); t_build = time.perf_counter() - t0
        res = {}

        def run():
            res["r"] = g.search(qh, k=10, ef_search=128, with_work=True)
        dt = timed(run)
        work = res["r"][3]
        evals = float(work[:, 0].mean())
        out["hnsw"] = {"n": hn, "dim": d, "M": 16, "ef": 128, "k": 10, "device_build_s": t_build, "qps_host_api": nq / dt,
                       "ms_per_batch": dt * 1e3, "dist_evals_per_query": evals, "expansions_per_query": float(work[:, 1].mean()),
                       "algorithmic_GBps": nq * evals * (d * 4 + 4) / dt / 1e9}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
