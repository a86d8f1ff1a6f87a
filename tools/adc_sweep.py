#!/usr/bin/env python
"""ADC scan tuning on the BASELINE configs[2] shape (IVFPQ, 768-d, nlist 4096, nprobe 32, M 96): builds the index once,
then times the search under different environment switches (read per call by the library)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from comet_b200 import capi  # noqa: E402

if os.environ.get("COMET_B200_SO"):          # A/B runs: another build of the library
    capi.SO_PATH = os.environ["COMET_B200_SO"]


def main():
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    ap.add_argument("--nlist", type=int, default=4096)
    ap.add_argument("--nq", type=int, default=512)
    ap.add_argument("--configs", default=";SLICES=2")
    ap.add_argument("--cuda-profiler", action="store_true", help="cudaProfilerStart() before the timed searches (ncu --profile-from-start off)")
    args = ap.parse_args()
    d, M, nprobe = 768, 96, 32
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(20261017)
    W3 = torch.randn((32, d), generator=g, device=dev)

    def rows(m):
        z = torch.randn((m, 32), generator=g, device=dev)
        return (z @ W3 + 0.05 * torch.randn((m, d), generator=g, device=dev)).cpu().numpy()

    ix = capi.IVFPQIndex(d, capi.L2, args.nlist, M, 8)
    ix.train(rows(args.nlist * 16))
    for s0 in range(0, args.n, 500_000):
        m = min(500_000, args.n - s0)
        ix.add(np.arange(s0 + 1, s0 + m + 1, dtype=np.uint32), rows(m), writeback=False)
    q = rows(args.nq)
    L = capi.lib()
    ref = None
    out = {}
    for cfg in args.configs.split(";"):
        for k in ("COMET_B200_ADC_SLICES", "COMET_B200_ADC_GENERIC", "COMET_B200_ADC_RING"):
            os.environ.pop(k, None)
        for kv in cfg.split(","):
            if kv:
                k, v = kv.split("=")
                os.environ["COMET_B200_ADC_" + k] = v
        r = ix.search(q, k=100, nprobes=nprobe)
        if ref is None:
            ref = r
        same = bool(np.array_equal(r[0], ref[0]) and np.array_equal(r[1].view(np.uint32), ref[1].view(np.uint32)))
        if args.cuda_profiler:
            torch.cuda.cudart().cudaProfilerStart()
        L.cm_profile_reset(); L.cm_profile_enable(1)
        t0 = time.perf_counter()
        for _ in range(5):
            ix.search(q, k=100, nprobes=nprobe)
        dt = (time.perf_counter() - t0) / 5
        L.cm_profile_enable(0)
        ms, cnt = capi.profile_get(capi.PROF_PQ_SCAN)
        coarse_ms = capi.profile_get(capi.PROF_FLAT_SCAN)[0] / 5
        select_ms = capi.profile_get(capi.PROF_SELECT)[0] / 5
        scanned = L.cm_ivfpq_last_scanned(ix.h) / args.nq
        out[cfg] = {"ms_per_batch": dt * 1e3, "qps": args.nq / dt, "adc_ms_per_batch": ms / 5, "launches_per_batch": cnt / 5,
                    "coarse_scan_ms": coarse_ms, "merges_ms": select_ms,
                    "lookups_per_s": args.nq * scanned * M / (ms / 5 * 1e-3), "same_bits_as_first": same}
        print(cfg, json.dumps(out[cfg]), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
