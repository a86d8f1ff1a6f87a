#!/usr/bin/env python
"""Flat 1M x 768 (cosine, K = 100): host-API time per call against the batch size, exact scan vs tensor path, with and
without a document filter -- the data behind CM_PATH_AUTO's thresholds."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from comet_b200 import capi  # noqa: E402


def main():
    import torch
    n, d, k = 1_000_000, 768, 100
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(1)
    ix = capi.FlatIndex(d, capi.COSINE)
    ix.reserve(n)
    x = torch.randn((n, d), generator=g, device=dev)
    ix.add_device(np.arange(1, n + 1, dtype=np.uint32), x.data_ptr(), n)
    torch.cuda.synchronize()
    del x
    rng = np.random.default_rng(2)
    q = rng.standard_normal((512, d), dtype=np.float32)
    filters = {"none": None, "half": np.arange(1, n + 1, 2, dtype=np.uint32), "1pct": np.arange(1, n + 1, 100, dtype=np.uint32),
               "300": rng.choice(n, 300, replace=False).astype(np.uint32) + 1}
    out = {}
    for fname, f in filters.items():
        for nq in (1, 4, 8, 16, 32, 64, 128, 256, 512):
            row = {}
            for pname, path in (("exact", capi.PATH_EXACT), ("tensor", capi.PATH_TENSOR), ("auto", capi.PATH_AUTO)):
                if pname == "exact" and nq > 64 and fname == "none":
                    continue
                try:
                    r0 = ix.search(q[:nq], k=k, path=path, filter_ids=f)
                    t0 = time.perf_counter()
                    reps = 5
                    for _ in range(reps):
                        r = ix.search(q[:nq], k=k, path=path, filter_ids=f)
                    row[pname] = round((time.perf_counter() - t0) / reps * 1e3, 3)
                    row[pname + "_path"] = ix.last_stats()["path_used"]
                except capi.CometError as e:
                    row[pname] = "err: " + e.msg[:60]
            out[f"{fname}/nq{nq}"] = row
            print(fname, nq, json.dumps(row), flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
