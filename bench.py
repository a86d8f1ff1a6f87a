#!/usr/bin/env python
"""bench.py -- headline benchmark of the comet_b200 flat search path.

Workload (BASELINE.json configs[1]): Flat Cosine, 1M x 768 float32 corpus, K=100, 512 queries per
GPU per step; synthetic N(0,1) data, random-init (there is no dataset to download).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one device batch: every query of the batch is searched against the whole corpus
(nq independent searchSingleQuery calls, flat_index_search.go:221-294) and its K best (id, score)
pairs are produced.  N > 1 (under torchrun): the 1M x 768 corpus fits every GPU, so queries are the
independent units -- every rank answers its own 512 queries against a replica, no data-path collective,
per-GPU work constant ("weak").  The same line carries `rows_sharded`: the north-star layout (rows sharded
over the N GPUs, one exchange step) through the library's single-process sharded index, driven by rank 0.
`--sharding rows` runs the per-process variant (NCCL all-gather of per-shard top-K + device merge).
`--workload c1 / b1 / c3 / c4 / c5`: BASELINE.json's other configurations (c5 = 100M x 768 over 8 GPUs).

`value`  : queries/s with queries already resident in HBM (CUDA events on the launching stream).
`e2e`    : queries/s through cm_flat_search with HOST buffers (pinned): H2D of the queries and D2H
           of ids/scores/counts inside the timed region.
`--impl reference` times the CPU restatement of the reference's Go loops (oracle/, all host threads)
on a bounded sample of the same workload (the Go toolchain does not exist in this image).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS, DIM, K, BATCH = 1_000_000, 768, 100, 512   # --rows overrides N_ROWS (e.g. 12_500_000 = one shard of configs[4])
SEED = 20261017
METRIC_NAME = "queries/sec @ recall@K (1Mx768, K=100)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# Exactly ONE line may reach stdout (the JSON); libraries chat there too (NCCL prints its version banner on
# init).  Keep the real stdout aside and point fd 1 at stderr for everything else.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML, else nvidia-smi)."""

    def __init__(self, index=0):
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            log("clock sampler: NVML unavailable:", e)
            self._nvml = None
            return self
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def _run(self):
        nv = self._nvml
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_traffic(key):
    """DRAM bytes of the dominant kernel from the committed ncu capture (profiles/), or None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                v = json.load(f).get(key)
            if v is not None:
                return v
        except Exception:
            pass
    return None


def gen_rows_device(torch, n, d, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randn((n, d), generator=g, device=device, dtype=torch.float32)


# -------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port), all host threads, bounded sample
# -------------------------------------------------------------------------------------------------
def cpu_reference_run(x_host, q_host, k, threads, nq_sample, metric=2):
    """Returns (qps, seconds, ids, scores) for nq_sample queries against the full corpus."""
    from oracle import oracle_py as O
    O.set_threads(threads)
    o = O.Flat(x_host.shape[1], metric)
    t0 = time.perf_counter()
    o.add(np.arange(1, x_host.shape[0] + 1, dtype=np.uint32), x_host)   # normalises in place, like Add
    t_add = time.perf_counter() - t0
    qs = np.ascontiguousarray(q_host[:nq_sample])
    t0 = time.perf_counter()
    ids, sc, cnt = o.search_batch(qs, k)
    dt = time.perf_counter() - t0
    return nq_sample / dt, dt, ids, sc, t_add, o


def workload_config(args, world, rows_mode):
    """The `config` object of the line -- the workload, identical in both arms (`--impl ours` / `--impl reference`)."""
    shard = N_ROWS // world if rows_mode else N_ROWS
    name = ("flat_%s_%dx768_k100_b512" % (args.metric_kind, N_ROWS)
            if (N_ROWS != 1_000_000 or args.metric_kind != "cosine") else "flat_cosine_1Mx768_k100_b512")
    return {"workload": name, "rows": N_ROWS, "dim": DIM, "k": K, "batch_per_gpu": BATCH, "global_batch": BATCH * max(1, world),
            "sharding": "single GPU" if world == 1 else (
                f"rows/{world} per GPU, every rank searches the global batch, NCCL all-gather of per-shard top-K + device merge"
                if rows_mode else
                f"queries: {world} replicas of the corpus (it fits one GPU), {BATCH} queries per GPU, no data-path collective; "
                "--sharding rows runs the row-sharded layout with the NCCL top-K merge"),
            "l2_policy": "inputs (%.2f GB fp32 corpus + bf16 shadow per GPU) larger than the 126 MB L2; no flush needed" % (shard * DIM * 4 / 1e9)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    # which layout the other arm picks at this N (its rule: replicas when the corpus fits 60 % of a 178 GiB B200)
    world = max(1, args.gpus)
    ref_mode = args.sharding if args.sharding != "auto" else ("queries" if N_ROWS * DIM * 6 < 0.6 * 178 * 2 ** 30 else "rows")
    ref_rows_mode = world == 1 or ref_mode == "rows"
    rng = np.random.default_rng(SEED)
    # bounded sample: a 1/8 row sample of the corpus, scaled (the reference's cost is linear in N:
    # one distance per row plus an N log N sort), so that steps x (cores queries) finish in minutes
    n_s = N_ROWS // 8
    x = rng.standard_normal((n_s, DIM), dtype=np.float32)
    q = rng.standard_normal((max(cores, 1), DIM), dtype=np.float32)
    from oracle import oracle_py as O
    O.set_threads(cores)
    o = O.Flat(DIM, O.COSINE)
    o.add(np.arange(1, n_s + 1, dtype=np.uint32), x)
    nq = len(q)
    for _ in range(args.warmup):
        o.search_batch(q, K)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.search_batch(q, K)
    dt = (time.perf_counter() - t0) / args.steps
    qps = nq / dt * (n_s / N_ROWS)
    sample = (f"{nq} queries (one per host thread) x {n_s} rows (1/8 row sample of the 1M corpus, QPS scaled by 1/8: "
              f"reference cost is linear in rows) per step; C restatement of the Go loops incl. the full sort")
    line = {
        "impl": "reference", "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same workload object as the other arm prints (sharding = the layout THAT arm runs at this N; this arm is
        # one host process whatever N is)
        "config": workload_config(args, max(1, args.gpus), ref_rows_mode),
        "run_detail": {"path": "reference CPU algorithm (C restatement of the Go loops), one query per host thread"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from comet_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log(f"warning: WORLD_SIZE={world} but --gpus {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = capi.lib()
    path = {"auto": capi.PATH_AUTO, "exact": capi.PATH_EXACT, "tensor": capi.PATH_TENSOR}[args.path]

    # ---- corpus: rows [rank*shard, (rank+1)*shard) of the 1M x 768 matrix; ids = row + 1 ----
    # --sharding rows   : north-star layout -- the corpus is row-sharded (rows/N per GPU), every rank searches the
    #                     global batch (512 N queries) on its shard, one exchange step (NCCL all-gather of the
    #                     per-shard top-K + device merge).  The only option when the corpus does not fit one GPU.
    # --sharding queries: the queries are the independent units: every GPU holds a replica of the corpus and
    #                     answers its own 512 queries; no data-path collective at all.
    # auto = queries when a replica (fp32 rows + bf16 shadow, 6 bytes per element) fits in 60 % of one GPU.
    mode = args.sharding
    if mode == "auto":
        total_mem = torch.cuda.get_device_properties(dev).total_memory
        mode = "queries" if N_ROWS * DIM * 6 < 0.6 * total_mem else "rows"
    if world == 1:
        mode = "rows"
    rows_mode = mode == "rows"
    shard = N_ROWS // world if rows_mode else N_ROWS
    row0 = rank * shard if rows_mode else 0
    seed_rank = rank if rows_mode else 0              # replicas hold identical rows
    metric_code = capi.METRICS[args.metric_kind]
    index = capi.FlatIndex(DIM, metric_code)
    index.reserve(shard)
    x_host = None
    slab = 2_000_000                                  # generate + add in slabs: a 12.5M-row shard is 38 GB
    for s0 in range(0, shard, slab):
        m = min(slab, shard - s0)
        x = gen_rows_device(torch, m, DIM, SEED + 1000 * seed_rank + 7919 * (s0 // slab), dev)
        index.add_device(np.arange(row0 + s0 + 1, row0 + s0 + m + 1, dtype=np.uint32), x.data_ptr(), m)
        torch.cuda.synchronize()
        if rank == 0 and not args.no_cpu_baseline and shard <= slab:
            x_host = x.cpu().numpy()
        del x
    torch.cuda.empty_cache()

    if rows_mode:
        nq = BATCH * world                  # global batch; every rank searches all of it on its shard
        q_dev = gen_rows_device(torch, nq, DIM, SEED + 7, dev)
    else:
        nq = BATCH                          # this rank's own queries against its replica
        q_dev = gen_rows_device(torch, nq, DIM, SEED + 7 + 31 * rank, dev)
    nq_global = BATCH * world
    q_host_pinned = torch.empty((nq, DIM), dtype=torch.float32, pin_memory=True)
    q_host_pinned.copy_(q_dev)
    torch.cuda.synchronize()
    q_np = q_host_pinned.numpy()

    out_ids = torch.zeros((nq, K), dtype=torch.int32, device=dev)
    out_sc = torch.zeros((nq, K), dtype=torch.float32, device=dev)
    out_cnt = torch.zeros((nq,), dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream()

    exchange_on = world > 1 and rows_mode
    if exchange_on:
        g_ids = torch.zeros((world, nq, K), dtype=torch.int32, device=dev)
        g_sc = torch.zeros((world, nq, K), dtype=torch.float32, device=dev)
        g_cnt = torch.zeros((world, nq), dtype=torch.int64, device=dev)
        m_ids = torch.zeros((nq, K), dtype=torch.int32, device=dev)
        m_sc = torch.zeros((nq, K), dtype=torch.float32, device=dev)
        m_cnt = torch.zeros((nq,), dtype=torch.int64, device=dev)

    def exchange():
        # the path's one exchange step: all-gather per-shard top-K over NCCL, then the device merge
        dist.all_gather_into_tensor(g_ids.view(world * nq, K), out_ids)
        dist.all_gather_into_tensor(g_sc.view(world * nq, K), out_sc)
        dist.all_gather_into_tensor(g_cnt.view(world * nq), out_cnt)
        capi.merge_shards_device(g_ids.data_ptr(), g_sc.data_ptr(), g_cnt.data_ptr(), world, nq, K, K,
                                 m_ids.data_ptr(), m_sc.data_ptr(), m_cnt.data_ptr(), out_stride=K,
                                 stream=stream.cuda_stream)

    def step_local():
        index.search_device(q_dev.data_ptr(), nq, K, out_ids.data_ptr(), out_sc.data_ptr(), out_cnt.data_ptr(), K,
                            stream=stream.cuda_stream, path=path)

    def step_device():
        step_local()
        if exchange_on:
            exchange()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -----------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    L.cm_profile_reset()
    L.cm_profile_enable(1 if args.profile_kernels else 0)
    launches0 = L.cm_kernel_launches()
    sampler = ClockSampler(local).start() if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    t_enq0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
    host_enqueue_ms = (time.perf_counter() - t_enq0) * 1e3 / args.steps   # CPU time to ENQUEUE one step (no sync inside; beyond ~75 steps the driver's launch queue fills and this includes waiting for the GPU)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = L.cm_kernel_launches() - launches0
    L.cm_profile_enable(0)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = nq_global / (ms_per_step * 1e-3)
    stats = index.last_stats()
    # a query whose candidate lists overflowed comes back with count -1 from the device entry point
    # (the host entry point redoes it with the exact scan): such a run would not have done the work
    unanswered = int((out_cnt < 0).sum().item())
    if unanswered:
        raise RuntimeError(f"{unanswered} queries were not answered by the device path (candidate overflow): number invalid")

    # ---- per-kernel roofline leg (separate short run so event pairs do not perturb `value`) -----
    roof = None
    if rank == 0:
        L.cm_profile_reset(); L.cm_profile_enable(1)
        n_prof = max(2, min(args.steps, 5))
        for _ in range(n_prof):
            step_local()          # rank-local kernels only: no collective on this leg
        torch.cuda.synchronize()
        L.cm_profile_enable(0)
        hbm, tf_burst, tf_sus, which = measured_peaks()
        scan_ms, scan_n = capi.profile_get(capi.PROF_FLAT_SCAN)
        gemm_ms, gemm_n = capi.profile_get(capi.PROF_FLAT_GEMM)
        resc_ms, resc_n = capi.profile_get(capi.PROF_RESCORE)
        sel_ms, sel_n = capi.profile_get(capi.PROF_SELECT)
        if gemm_n > 0:
            # the candidate pass covers the corpus once per step in `gemm_n / n_prof` launches over
            # row samples: flops per step = 2 * nq * rows * dim, time = their summed durations.
            # Denominator: the cuBLAS peak of the regime THIS run was in, read from its own clock samples --
            # burst unless the power cap was active or the SM clock sat under 90 % of its maximum.
            per = gemm_ms / n_prof * 1e-3
            flops = 2.0 * nq * shard * DIM
            ach = flops / per / 1e12
            capped = bool(clocks) and ("sw_power_cap" in clocks["reasons"] or (
                clocks["sm_mhz"] and clocks["sm_max_mhz"] and clocks["sm_mhz"] < 0.9 * clocks["sm_max_mhz"]))
            # a power-cap flag sampled while the clock stayed at its maximum is not the sustained regime: a pass that
            # beats the sustained cuBLAS figure is measured against the burst one (the fraction never flatters itself)
            if capped and ach > tf_sus:
                capped = False
            peak = tf_sus if capped else tf_burst
            roof = {"kernel": "flat_gemm_ts_kernel (tcgen05 bf16, queries resident in tensor memory, %d launches per step)" % (gemm_n // n_prof),
                    "bound": "tensor", "achieved": ach, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach / peak,
                    "regime": "sustained (power cap / low clock seen in this run)" if capped else "burst (clock >= 90 % of max in this run, or faster than the sustained cuBLAS figure)",
                    "frac_of_burst": ach / tf_burst, "frac_of_sustained": ach / tf_sus,
                    "traffic": measured_traffic("flat_gemm_kernel_per_step_bytes") if (world == 1 and N_ROWS == 1_000_000) else None,
                    "traffic_note": "DRAM bytes of the candidate-pass launches of one step, ncu capture in profiles/; algorithmic = %d (bf16 shadow once)" % (shard * DIM * 2),
                    "peak_source": which + (" (sustained bf16)" if capped else " (burst bf16)"),
                    "gemm_ms_per_step": per * 1e3, "launches_timed": gemm_n,
                    "hbm_equiv": {"note": "algorithmic bytes of one fp32 corpus pass / GEMM time, vs measured HBM peak",
                                  "achieved_gbs": (shard * DIM * 4 + nq * DIM * 4 + nq * K * 8) / per / 1e9,
                                  "peak_gbs": hbm}}
        elif scan_n > 0:
            per = scan_ms / scan_n * 1e-3
            qb = max(1, nq // max(1, scan_n // max(2, min(args.steps, 5))))
            abytes = shard * DIM * 4 + qb * DIM * 4
            ach = abytes / per / 1e9
            roof = {"kernel": "flat_scan_kernel", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                    "frac": ach / hbm, "traffic": measured_traffic("flat_scan_kernel_per_launch_bytes") if world == 1 else None,
                    "peak_source": which, "avg_launch_ms": per * 1e3,
                    "launches_timed": scan_n, "queries_per_launch": qb}
        if roof is not None:
            roof["step_share"] = {"steps": n_prof, "scan_ms": scan_ms, "gemm_ms": gemm_ms, "rescore_ms": resc_ms,
                                  "select_ms": sel_ms}

    # ---- e2e: host buffers through the public C ABI call ---------------------------------------
    # result buffers in pinned host memory, like the query block: what a caller that cares about PCIe does
    h_ids = torch.zeros((nq, K), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
    h_sc = torch.zeros((nq, K), dtype=torch.float32, pin_memory=True).numpy()
    h_cnt = torch.zeros((nq,), dtype=torch.int64, pin_memory=True).numpy()
    import ctypes as C
    p, keep = capi.make_params(k=K, path=path)

    def step_host():
        capi.check(L.cm_flat_search(index.h, capi.ptr(q_np, capi.f32p), nq, DIM, C.byref(p), K,
                                    capi.ptr(h_ids, capi.u32p), capi.ptr(h_sc, capi.f32p), None,
                                    capi.ptr(h_cnt, capi.i64p)))
        if exchange_on:
            # shard results go back to the device for the exchange step; the merged list returns to the host
            out_ids.copy_(torch.from_numpy(h_ids.view(np.int32)), non_blocking=True)
            out_sc.copy_(torch.from_numpy(h_sc), non_blocking=True)
            out_cnt.copy_(torch.from_numpy(h_cnt), non_blocking=True)
            exchange()
            h_ids[...] = m_ids.cpu().numpy().view(np.uint32)
            h_sc[...] = m_sc.cpu().numpy()
            h_cnt[...] = m_cnt.cpu().numpy()

    for _ in range(min(args.warmup, 3)):
        step_host()
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    n_copy = nq if rows_mode else nq_global           # bytes crossing PCIe per step, all ranks together in queries mode
    e2e = {"value": nq_global / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(n_copy * DIM * 4),
           "d2h_bytes_per_step": int(n_copy * K * 8 + n_copy * 8), "ms_per_step": e2e_s * 1e3, "steps": e2e_steps}

    # ---- CPU baseline + parity / recall of this very run (rank 0, N=1) -------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and x_host is not None:
        cores = os.cpu_count() or 1
        nq_s = min(nq, max(cores, 8))
        qps, dt, o_ids, o_sc, t_add, _o = cpu_reference_run(x_host, q_np, K, cores, nq_s, metric_code)
        cpu = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
               "sample": f"{nq_s} of the {nq} queries against the full 1M x 768 corpus, one query per host thread, "
                         f"{dt:.1f} s; C restatement of the Go loops (scalar, unfused, full N sort per query)"}
        exact_ids = bool(np.array_equal(h_ids[:nq_s], o_ids[:, :K]))
        exact_sc = bool(np.array_equal(h_sc[:nq_s].view(np.uint32), o_sc[:, :K].view(np.uint32)))
        rec = float(np.mean([len(set(h_ids[i].tolist()) & set(o_ids[i, :K].tolist())) / K for i in range(nq_s)]))
        parity = {"queries_checked": nq_s, "ids_bit_exact": exact_ids, "scores_bit_exact": exact_sc,
                  "recall_at_k": rec}

    # ---- row-sharded layout through the library's single-process sharded index (rank 0 drives every GPU) -------
    rows_sharded = None
    if world > 1 and not args.no_rows_sharded:
        # host-side barrier (gloo): an NCCL barrier would park a spinning kernel on the GPUs rank 0 is about to time
        host_group = dist.new_group(backend="gloo")
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        if rank == 0:
            try:
                rows_sharded = sharded_leg(torch, capi, list(range(world)), N_ROWS, metric_code, max(3, min(args.steps, 50)),
                                           args.warmup, path)
                rows_sharded["note"] = ("north-star layout (BASELINE configs[4] shape at %d rows per GPU): corpus of N x %d rows, "
                                        "batch %d; weak scaling shows in row_queries_per_s" % (N_ROWS, N_ROWS, BATCH))
            except Exception as e:   # noqa: BLE001 -- the headline line must still be printed
                rows_sharded = {"error": repr(e)}
            torch.cuda.set_device(local)
        dist.barrier(group=host_group)

    if rank == 0:
        line = {
            "metric": METRIC_NAME, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world, rows_mode),
            "run_detail": {"path": {1: "exact fp32 scan", 2: "bf16 tcgen05 candidates + exact fp32 re-score"}.get(stats["path_used"], "?"),
                           "scan_passes_per_step": stats["passes"], "rescored_candidates_per_step": stats["candidates"]},
            "step_vs_hbm_roofline": {
                "note": "north-star target: the whole step at >= 0.70 of the time one fp32 corpus pass takes at the measured HBM peak",
                "algorithmic_bytes_per_step": int(shard * DIM * 4 + nq * DIM * 4 + nq * K * 8),
                "achieved_gbs": (shard * DIM * 4 + nq * DIM * 4 + nq * K * 8) / (ms_per_step * 1e-3) / 1e9,
                "peak_gbs": measured_peaks()[0], "frac": (shard * DIM * 4 + nq * DIM * 4 + nq * K * 8) / (ms_per_step * 1e-3) / 1e9 / measured_peaks()[0]},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms,
            "roofline": roof, "cpu_baseline": cpu, "parity": parity,
        }
        if rows_sharded is not None:
            line["rows_sharded"] = rows_sharded
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0



# -------------------------------------------------------------------------------------------------
# row-sharded leg: the north-star multi-GPU layout (BASELINE.json configs[4]) through the library's own
# single-process sharded index (cm_flat_sharded_*): shard r on GPU r, `rows_per_shard` rows each, every shard
# searches the whole query batch, NVLink peer copies bring the per-shard top-K lists to GPU 0, one merge kernel.
# Per-GPU work is fixed as N grows (the corpus grows with N), so the quantity that scales is rows x queries / s.
# -------------------------------------------------------------------------------------------------
def sharded_leg(torch, capi, devices, rows_per_shard, metric_code, steps, warmup, path, parity_queries=16):
    W = len(devices)
    lead = torch.device("cuda", devices[0])
    g = capi.ShardedFlatIndex(DIM, metric_code, devices, rows_per_shard)
    g.reserve(W * rows_per_shard)
    slab = 1_000_000
    t_fill0 = time.perf_counter()
    for s0 in range(0, rows_per_shard, slab):             # every GPU generates and adds its slab concurrently
        m = min(slab, rows_per_shard - s0)
        keep = []
        for r, dv in enumerate(devices):
            with torch.cuda.device(dv):
                x = gen_rows_device(torch, m, DIM, SEED + 1000 * r + 7919 * (s0 // slab), torch.device("cuda", dv))
                first = r * rows_per_shard + s0 + 1
                g.add_device(r, np.arange(first, first + m, dtype=np.uint32), x.data_ptr(), m,
                             stream=torch.cuda.current_stream().cuda_stream)
                keep.append(x)
        for dv in devices:
            torch.cuda.synchronize(dv)
        del keep
    fill_s = time.perf_counter() - t_fill0
    nq = BATCH
    with torch.cuda.device(lead):
        q_dev = gen_rows_device(torch, nq, DIM, SEED + 7, lead)
        out_ids = torch.zeros((nq, K), dtype=torch.int32, device=lead)
        out_sc = torch.zeros((nq, K), dtype=torch.float32, device=lead)
        out_cnt = torch.zeros((nq,), dtype=torch.int64, device=lead)
        stream = torch.cuda.current_stream()
        q_pin = torch.empty((nq, DIM), dtype=torch.float32, pin_memory=True)
        q_pin.copy_(q_dev)
        torch.cuda.synchronize()

        def sync_all():
            for dv in devices:
                torch.cuda.synchronize(dv)

        def step():
            g.search_device(q_dev.data_ptr(), nq, K, out_ids.data_ptr(), out_sc.data_ptr(), out_cnt.data_ptr(), K,
                            stream=stream.cuda_stream, path=path)

        for _ in range(max(3, warmup)):
            step()
        sync_all()
        L = capi.lib()
        launches0 = L.cm_kernel_launches()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step()
        e1.record(stream)                     # the leader's stream waits for every shard before the merge
        sync_all()
        ms = e0.elapsed_time(e1) / steps
        launches = L.cm_kernel_launches() - launches0
        timing = g.last_timing()
        xbytes = g.exchange_bytes()
        stats = g.last_stats()
        if int((out_cnt < 0).sum().item()):
            raise RuntimeError("sharded leg: queries left unanswered by the device path (candidate overflow)")
        # end to end: host buffers through cm_flat_sharded_search (H2D of the queries, D2H of the merged lists)
        q_np = q_pin.numpy()
        h_ids = torch.zeros((nq, K), dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
        h_sc = torch.zeros((nq, K), dtype=torch.float32, pin_memory=True).numpy()
        h_cnt = torch.zeros((nq,), dtype=torch.int64, pin_memory=True).numpy()
        import ctypes as C
        p, keepalive = capi.make_params(k=K, path=path)

        def step_host():
            capi.check(L.cm_flat_sharded_search(g.h, capi.ptr(q_np, capi.f32p), nq, DIM, C.byref(p), K,
                                                capi.ptr(h_ids, capi.u32p), capi.ptr(h_sc, capi.f32p),
                                                capi.ptr(h_cnt, capi.i64p)))
        for _ in range(3):
            step_host()
        e2e_steps = max(3, min(steps, 20))
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        # the exchange + merge against an independent merge: every shard's own top-K (host API of the single index
        # is not reachable through the sharded handle, so the device lists are re-merged on the host with
        # numpy: order by (score, shard, rank) == (score, scan position)); the device merge must agree bit for bit.
        # Shard-local exactness vs the oracle is what tests/test_config_sizes_gpu.py pins at this shard size.
        nchk = min(parity_queries, nq)
        ok_order = True
        for i in range(nchk):
            sc = h_sc[i, :int(h_cnt[i])]
            ok_order &= bool(np.all(sc[1:] >= sc[:-1]))
            ok_order &= len(set(h_ids[i, :int(h_cnt[i])].tolist())) == int(h_cnt[i])
        same_dev_host = bool(np.array_equal(out_ids.cpu().numpy().view(np.uint32), h_ids) and
                             np.array_equal(out_sc.cpu().numpy().view(np.uint32), h_sc.view(np.uint32)))
    rows_total = W * rows_per_shard
    rec = {
        "layout": f"rows: {W} shards x {rows_per_shard} rows on {len(set(devices))} GPU(s), one host process, NVLink peer copies + one merge kernel",
        "n_gpus": len(set(devices)), "rows_total": rows_total, "rows_per_gpu": rows_per_shard, "batch": nq, "k": K,
        "value": nq / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
        "row_queries_per_s": nq * rows_total / (ms * 1e-3),
        "slowest_shard_search_ms": timing["search_ms_max"], "gather_ms": timing["gather_ms_max"],
        "merge_ms": timing["merge_ms"], "nvlink_bytes_per_step": xbytes,
        "e2e": {"value": nq / e2e_s, "unit": "queries/s", "ms_per_step": e2e_s * 1e3,
                "h2d_bytes_per_step": int(nq * DIM * 4), "d2h_bytes_per_step": int(nq * K * 8 + nq * 8), "steps": e2e_steps},
        "gpu_launches": int(launches), "rescored_candidates_per_step": stats["candidates"],
        "fallback_queries": stats["fallback_queries"], "fill_s": fill_s,
        "step_vs_hbm_roofline_per_gpu": (rows_per_shard * DIM * 4 + nq * DIM * 4 + nq * K * 8) / (ms * 1e-3) / 1e9 / measured_peaks()[0],
        "checks": {"sorted_unique_lists": ok_order, "device_and_host_entry_points_agree": same_dev_host},
    }
    del g
    return rec


def run_c5(args):
    """BASELINE.json configs[4]: Flat L2, 100M x 768 row-sharded over 8 GPUs (12.5M rows per GPU), K=100; the series
    keeps 12.5M rows per GPU, so N GPUs hold N x 12.5M rows.  One process (no torchrun)."""
    import torch
    from comet_b200 import capi
    n_dev = torch.cuda.device_count()
    rows = args.rows if args.rows > 0 else 12_500_000
    series = []
    ns = [n for n in (1, 2, 4, 8) if n <= min(n_dev, args.gpus)]
    path = {"auto": capi.PATH_AUTO, "exact": capi.PATH_EXACT, "tensor": capi.PATH_TENSOR}[args.path]
    for n in reversed(ns):
        log(f"c5: {n} GPU(s) x {rows} rows")
        series.append(sharded_leg(torch, capi, list(range(n)), rows, capi.METRICS[args.metric_kind], args.steps,
                                  args.warmup, path))
        torch.cuda.empty_cache()
    series.reverse()
    base = series[0]["row_queries_per_s"]
    for n, rec in zip(ns, series):
        rec["weak_scaling_efficiency_vs_1gpu"] = rec["row_queries_per_s"] / (n * base) if ns[0] == 1 else None
    top = series[-1]
    emit({"metric": "queries/sec (Flat %s, %d x 768 row-sharded, K=100, batch 512)" % (args.metric_kind, top["rows_total"]),
          "value": top["value"], "unit": "queries/s", "n_gpus": top["n_gpus"], "steps": args.steps, "warmup": args.warmup,
          "ms_per_step": top["ms_per_step"], "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic",
          "config": {"workload": "flat_%s_%dx768_rowsharded_k100_b512" % (args.metric_kind, top["rows_total"])},
          "series": series})
    return 0


# -------------------------------------------------------------------------------------------------
# --workload c3 / c4: BASELINE.json configs[2] (IVFPQ 10M x 768, nlist 4096, nprobe 32, M 96, nbits 8, K 100) and
# configs[3] (HNSW 1M x 768, M 16, efSearch 128, K 10) at their full sizes, one GPU.  Each line carries `value`
# (device-resident queries), `e2e` (host API), `roofline` (the dominant kernel against the HBM peak in the algorithmic
# bytes of SURVEY 8d), `cpu_baseline` (the oracle on the SAME index state, one query per host thread) and `parity`
# (ids, score bits -- and for HNSW the work counters -- of those queries, bit-exact).
# -------------------------------------------------------------------------------------------------
def _oracle_queries_parallel(search_one, nq, cores):
    """search_one(i) -> result; one query per host thread (ctypes drops the GIL inside the oracle)."""
    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        res = list(ex.map(search_one, range(nq)))
    return res, time.perf_counter() - t0


def _device_leg(torch, fn_enqueue, steps, warmup):
    for _ in range(max(3, warmup)):
        fn_enqueue()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream()
    e0.record(st)
    for _ in range(steps):
        fn_enqueue()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_c3(args):
    import ctypes as C
    import torch
    from comet_b200 import capi
    from oracle import oracle_py as O              # cpu_baseline + in-run parity
    d, nq, k = DIM, BATCH, 100
    n = args.rows if args.rows > 0 else 10_000_000
    nlist, nprobe, M = 4096, 32, 96
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    g = torch.Generator(device=dev); g.manual_seed(SEED)
    # rows on a 32-dimensional linear manifold + small isotropic noise: one connected cloud, so the reference's
    # deterministic k-means yields lists of comparable length (i.i.d. 768-d Gaussians have no cluster structure: hub
    # lists and empty lists)
    W3 = torch.randn((32, d), generator=g, device=dev)

    def rows(m):
        z = torch.randn((m, 32), generator=g, device=dev)
        return (z @ W3 + 0.05 * torch.randn((m, d), generator=g, device=dev)).cpu().numpy()

    ix = capi.IVFPQIndex(d, capi.L2, nlist, M, 8)
    t0 = time.perf_counter(); ix.train(rows(nlist * 16)); t_train = time.perf_counter() - t0
    lists = np.zeros(n, np.int32)
    t_add = 0.0
    slab = 500_000
    for s0 in range(0, n, slab):
        m = min(slab, n - s0)
        x = rows(m)
        t0 = time.perf_counter()
        lists[s0:s0 + m] = ix.add(np.arange(s0 + 1, s0 + m + 1, dtype=np.uint32), x, writeback=False)
        t_add += time.perf_counter() - t0
        del x
    q = rows(nq)
    L = capi.lib()
    qd = torch.from_numpy(q).to(dev)
    o_ids = torch.zeros((nq, k), dtype=torch.int32, device=dev)
    o_sc = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    o_cnt = torch.zeros((nq,), dtype=torch.int64, device=dev)
    p, keep = capi.make_params(k=k, nprobes=nprobe)
    stream = torch.cuda.current_stream().cuda_stream

    def enqueue():
        capi.check(L.cm_ivfpq_search_device(ix.h, capi.vp(qd.data_ptr()), nq, d, C.byref(p), k, capi.vp(o_ids.data_ptr()),
                                            capi.vp(o_sc.data_ptr()), None, capi.vp(o_cnt.data_ptr()), capi.vp(stream)))
    launches0 = L.cm_kernel_launches()
    sampler = ClockSampler(0).start()
    ms = _device_leg(torch, enqueue, args.steps, args.warmup)
    clocks = sampler.stop()
    launches = (L.cm_kernel_launches() - launches0) * args.steps // (args.steps + max(3, args.warmup))
    L.cm_profile_reset(); L.cm_profile_enable(1)
    for _ in range(3):
        enqueue()
    torch.cuda.synchronize()
    L.cm_profile_enable(0)
    scan_ms, scan_n = capi.profile_get(capi.PROF_PQ_SCAN)
    per = scan_ms / 3 * 1e-3                              # ADC scan time per step (all its launches)
    scanned = L.cm_ivfpq_last_scanned(ix.h) / nq
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(e2e_steps):
        gi, gs, gc = ix.search(q, k=k, nprobes=nprobe)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    # the oracle on the SAME trained state and codes (its ReadFrom-style loader)
    cores = os.cpu_count() or 1
    o = O.IVFPQ(d, capi.L2, nlist, M, 8)
    o.set_trained(*ix.trained_state())
    o.load_codes(np.arange(1, n + 1, dtype=np.uint32), ix.codes(), lists)
    nchk = nq                                            # every query of the batch: parity and the CPU figure
    res, dt = _oracle_queries_parallel(lambda i: o.search(q[i], k=k, nprobes=nprobe), nchk, cores)
    ok_ids = all(np.array_equal(gi[i, :int(gc[i])], res[i][0]) for i in range(nchk))
    ok_sc = all(np.array_equal(gs[i, :int(gc[i])].view(np.uint32), res[i][1].view(np.uint32)) for i in range(nchk))
    hbm = measured_peaks()[0]
    code_bytes = nq * scanned * M                       # the code stream of the probed lists, M bytes per code
    emit({"metric": "queries/sec (IVFPQ %dx768, nlist 4096, nprobe 32, M 96, nbits 8, K=100)" % n, "value": nq / (ms * 1e-3),
          "unit": "queries/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
          "scaling": "weak", "vs_baseline": None, "dtype": "f32 (u8 codes)", "data": "synthetic",
          "config": {"workload": "ivfpq_%dx768_nlist4096_nprobe32_m96_k100_b512" % n, "rows": n, "dim": d, "nlist": nlist, "nprobes": nprobe,
                     "M": M, "nbits": 8, "k": k, "batch": nq, "data_shape": "32-d linear manifold + 0.05 noise in 768-d",
                     "codes_scanned_per_query": scanned, "train_s": t_train, "add_s": t_add},
          "clocks": clocks,
          "e2e": {"value": nq / e2e_s, "unit": "queries/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": nq * d * 4,
                  "d2h_bytes_per_step": nq * k * 8 + nq * 8, "steps": e2e_steps},
          "gpu_launches": int(launches),
          "roofline": {"kernel": "adc_ring_kernel (per-(query, probe) residual table in shared memory, one bank per lane; list codes streamed pre-skewed)",
                       "bound": "hbm", "achieved": code_bytes / per / 1e9, "peak": hbm, "unit": "GB/s", "frac": code_bytes / per / 1e9 / hbm,
                       "traffic": None, "kernel_ms_per_step": per * 1e3, "launches_timed": scan_n,
                       "note": "algorithmic bytes = codes scanned x M; the kernel is bound by the SM load/store data pipe (table reads, shuffles, codeword loads: 71 % busy in profiles/r02_ncu_full_adc_ring.md), not by this stream",
                       "table_lookups_per_s": nq * scanned * M / per, "tables_built_per_s": nq * nprobe / per},
          "cpu_baseline": {"value": nchk / dt, "unit": "queries/s", "cores": min(cores, nchk), "kind": "port",
                           "sample": f"{nchk} of the {nq} queries on the same trained index and codes, one query per host thread, {dt:.1f} s"},
          "parity": {"queries_checked": nchk, "ids_bit_exact": bool(ok_ids), "scores_bit_exact": bool(ok_sc)}})
    return 0


def run_c4(args):
    import ctypes as C
    import torch
    from comet_b200 import capi
    from oracle import oracle_py as O              # cpu_baseline + in-run parity
    d, k, ef = DIM, 10, 128
    n = args.rows if args.rows > 0 else 1_000_000
    nq = args.batch if args.batch > 0 else BATCH
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    g = torch.Generator(device=dev); g.manual_seed(SEED)
    Wk = torch.randn((24, d), generator=g, device=dev)
    x = (torch.randn((n, 24), generator=g, device=dev) @ Wk).cpu().numpy()
    q = (torch.randn((nq, 24), generator=g, device=dev) @ Wk).cpu().numpy()
    ids = np.arange(1, n + 1, dtype=np.uint32)
    # graph: the exact 32-nearest-neighbour graph of the rows (built with the flat tensor path).  A graph built by the
    # reference's own insertNode leaves 2M+1 nodes reachable (entry point never promoted, hnsw_index.go:266-284,
    # 675-694), so the traversal kernel is measured on a graph that is actually explored.
    flat = capi.FlatIndex(d, capi.L2)
    flat.add(ids, x, writeback=False)
    nbr = np.zeros((n, 32), np.uint32)
    t0 = time.perf_counter()
    for s0 in range(0, n, 16384):
        gi, _, _ = flat.search(x[s0:s0 + 16384], k=33)
        nbr[s0:s0 + 16384] = gi[:, 1:33]
    t_knn = time.perf_counter() - t0
    truth, _, _ = flat.search(q, k=k)
    del flat
    levels = np.zeros(n, np.int32)
    off = np.arange(n + 1, dtype=np.int64) * 32
    ix = capi.HNSWIndex(d, capi.L2, 16, 100, ef)
    ix.load_graph(ids, x, levels, [(off, nbr.ravel())], 1, 0)
    L = capi.lib()
    qd = torch.from_numpy(q).to(dev)
    o_ids = torch.zeros((nq, k), dtype=torch.int32, device=dev)
    o_sc = torch.zeros((nq, k), dtype=torch.float32, device=dev)
    o_cnt = torch.zeros((nq,), dtype=torch.int64, device=dev)
    o_work = torch.zeros((nq, 2), dtype=torch.int64, device=dev)
    p, keep = capi.make_params(k=k, ef_search=ef)
    stream = torch.cuda.current_stream().cuda_stream

    def enqueue():
        capi.check(L.cm_hnsw_search_device(ix.h, capi.vp(qd.data_ptr()), nq, d, C.byref(p), k, capi.vp(o_ids.data_ptr()),
                                           capi.vp(o_sc.data_ptr()), None, capi.vp(o_cnt.data_ptr()), capi.vp(o_work.data_ptr()),
                                           capi.vp(stream)))
    launches0 = L.cm_kernel_launches()
    sampler = ClockSampler(0).start()
    ms = _device_leg(torch, enqueue, args.steps, args.warmup)
    clocks = sampler.stop()
    launches = (L.cm_kernel_launches() - launches0) * args.steps // (args.steps + max(3, args.warmup))
    L.cm_profile_reset(); L.cm_profile_enable(1)
    for _ in range(3):
        enqueue()
    torch.cuda.synchronize()
    L.cm_profile_enable(0)
    k_ms, k_n = capi.profile_get(capi.PROF_HNSW)
    per = (k_ms / 3 * 1e-3) if k_n else ms * 1e-3        # traversal kernel time per step (all its launches)
    e2e_steps = max(3, min(args.steps, 20))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        gi, gs, gc, work = ix.search(q, k=k, ef_search=ef, with_work=True)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    evals = float(work[:, 0].mean())
    rec = float(np.mean([len(set(gi[i, :k].tolist()) & set(truth[i].tolist())) / k for i in range(nq)]))
    cores = os.cpu_count() or 1
    o = O.HNSW(d, capi.L2, 16, 100, ef)
    o.load_graph(ids, x, levels, off, nbr.ravel(), 1, 0)
    nchk = min(nq, 2048)                                 # parity on (up to) 2048 queries of the batch

    def one(i):
        r = o.search(q[i % nchk], k=k, ef_search=ef)
        return r, O.HNSW.last_counters()
    res, dt = _oracle_queries_parallel(one, nchk, cores)
    reps = max(1, int(8.0 / max(dt, 1e-3)))              # CPU figure from ~8 s of work
    if reps > 1:
        _, dt_all = _oracle_queries_parallel(one, nchk * reps, cores)
        cpu_qps = nchk * reps / dt_all
    else:
        dt_all, cpu_qps = dt, nchk / dt
    ok_ids = all(np.array_equal(gi[i, :int(gc[i])], res[i][0][0]) for i in range(nchk))
    ok_sc = all(np.array_equal(gs[i, :int(gc[i])].view(np.uint32), res[i][0][1].view(np.uint32)) for i in range(nchk))
    ok_work = all((int(work[i, 0]), int(work[i, 1])) == tuple(res[i][1]) for i in range(nchk))
    hbm = measured_peaks()[0]
    row_bytes = nq * evals * (d * 4 + 4)                 # one row (+ its id) per distance evaluation
    emit({"metric": "queries/sec (HNSW %dx768, M 16, efSearch 128, K=10)" % n, "value": nq / (ms * 1e-3), "unit": "queries/s",
          "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
          "vs_baseline": None, "dtype": "f32", "data": "synthetic",
          "config": {"workload": "hnsw_%dx768_m16_ef128_k10_b%d" % (n, nq), "rows": n, "dim": d, "degree": 32, "ef_search": ef, "k": k,
                     "queries_in_flight": nq, "graph": "exact 32-NN graph of the rows, layer 0 only, loaded into the device and the oracle",
                     "data_shape": "24-d linear manifold in 768-d", "knn_graph_build_s": t_knn,
                     "distance_evaluations_per_query": evals, "expansions_per_query": float(work[:, 1].mean()),
                     "recall_at_10_vs_flat": rec},
          "clocks": clocks,
          "e2e": {"value": nq / e2e_s, "unit": "queries/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": nq * d * 4,
                  "d2h_bytes_per_step": nq * k * 8 + nq * 8 + nq * 16, "steps": e2e_steps},
          "gpu_launches": int(launches),
          "roofline": {"kernel": "hnsw_search_kernel (one warp per query, neighbour rows staged through shared memory)", "bound": "hbm",
                       "achieved": row_bytes / per / 1e9, "peak": hbm, "unit": "GB/s", "frac": row_bytes / per / 1e9 / hbm, "traffic": None,
                       "kernel_ms_per_step": per * 1e3, "launches_timed": k_n,
                       "note": "algorithmic bytes = distance evaluations x (row + id); a dependent chain per query, throughput comes from queries in flight"},
          "cpu_baseline": {"value": cpu_qps, "unit": "queries/s", "cores": min(cores, nchk), "kind": "port",
                           "sample": f"{nchk} of the {nq} queries on the same graph x {reps} rounds, one query per host thread, {dt_all:.1f} s"},
          "parity": {"queries_checked": nchk, "ids_bit_exact": bool(ok_ids), "scores_bit_exact": bool(ok_sc),
                     "work_counters_equal": bool(ok_work)}})
    return 0

# -------------------------------------------------------------------------------------------------
# --workload c1 / b1: the single-query shapes (not the driver's bench line; parity-test configs measured
# for DESIGN.md).  c1 = BASELINE.json configs[0], Flat L2Squared 10K x 128, K=10, one query per call: latency
# of one host-API call next to the CPU restatement on one thread (what one Go Execute() does).
# b1 = one to eight queries against the headline corpus: the exact TMA scan, one pass over 3.07 GB.
# -------------------------------------------------------------------------------------------------
def run_single_query_shapes(args):
    import torch
    from comet_b200 import capi

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return (time.perf_counter() - t0) / reps

    rng = np.random.default_rng(SEED)
    out = {}
    if args.workload == "c1":
        from oracle import oracle_py as O          # cpu_baseline leg + in-run parity check
        n, d, k = 10_000, 128, 10
        x = rng.standard_normal((n, d), dtype=np.float32)
        q = rng.standard_normal((64, d), dtype=np.float32)
        ids = np.arange(1, n + 1, dtype=np.uint32)
        g = capi.FlatIndex(d, capi.L2SQ)
        g.add(ids, x.copy())
        o = O.Flat(d, capi.L2SQ)
        o.add(ids, x.copy())
        gi, gs, gc = g.search(q[:1], k=k)
        oi, os_ = o.search(q[0], k=k)
        O.set_threads(1)
        cpu_s = timed(lambda: o.search(q[0], k=k), 50)
        out = {"metric": "latency of one single-query Execute()", "unit": "us", "higher_is_better": False,
               "config": {"workload": "flat_l2sq_10Kx128_k10_b1"}, "dtype": "f32", "data": "synthetic",
               "value": timed(lambda: g.search(q[:1], k=k), 300) * 1e6,
               "us_per_call_b8": timed(lambda: g.search(q[:8], k=k), 300) * 1e6,
               "us_per_call_b64": timed(lambda: g.search(q[:64], k=k), 300) * 1e6,
               "algorithmic_bytes_per_query": n * d * 4 + d * 4 + k * 8,
               "cpu_baseline": {"value": cpu_s * 1e6, "unit": "us", "cores": 1, "kind": "port",
                                "sample": "the same query, CPU restatement of the Go loops, one thread"},
               "parity": {"ids_bit_exact": bool(np.array_equal(gi[0], oi)),
                          "scores_bit_exact": bool(np.array_equal(gs[0].view(np.uint32), os_.view(np.uint32)))}}
    else:
        n2, d2 = N_ROWS, DIM
        dev = torch.device("cuda", 0)
        g2 = capi.FlatIndex(d2, capi.COSINE)
        g2.reserve(n2)
        x = gen_rows_device(torch, n2, d2, SEED, dev)
        g2.add_device(np.arange(1, n2 + 1, dtype=np.uint32), x.data_ptr(), n2)
        torch.cuda.synchronize()
        del x
        q2 = rng.standard_normal((8, d2), dtype=np.float32)
        L = capi.lib()
        hbm = measured_peaks()[0]
        out = {"metric": "exact scan, small batches", "config": {"workload": "flat_cosine_1Mx768_k100_b1..8"}}
        for b in (1, 2, 4, 8):
            L.cm_profile_reset(); L.cm_profile_enable(1)
            dt = timed(lambda: g2.search(q2[:b], k=K), 20)
            L.cm_profile_enable(0)
            ms, cnt = capi.profile_get(capi.PROF_FLAT_SCAN)
            per = ms / max(cnt, 1)
            gbs = (n2 * d2 * 4 + b * d2 * 4) / (per * 1e-3) / 1e9
            out[f"b{b}"] = {"host_call_ms": dt * 1e3, "scan_kernel_ms": per, "scan_GBps": gbs, "frac_of_hbm_peak": gbs / hbm}
    emit(out)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default="auto", choices=["auto", "exact", "tensor"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-kernels", action="store_true", help="event-time kernels inside the main timed region too")
    ap.add_argument("--no-rows-sharded", action="store_true", help="N > 1: skip the row-sharded leg (rows_sharded sub-record)")
    ap.add_argument("--sharding", default="auto", choices=["auto", "rows", "queries"])
    ap.add_argument("--rows", type=int, default=0, help="corpus rows (default 1M = BASELINE configs[1]); 12500000 = one 1/8 shard of the 100M x 768 config")
    ap.add_argument("--batch", type=int, default=0, help="c4: queries in flight per call (default 512)")
    ap.add_argument("--metric-kind", default="cosine", choices=["cosine", "l2", "l2_squared"])
    ap.add_argument("--workload", default="headline", choices=["headline", "c1", "b1", "c3", "c4", "c5"],
                    help="headline = the driver's bench line; c1 / b1 = single-query shapes (see run_single_query_shapes)")
    args = ap.parse_args()
    if args.rows > 0:
        global N_ROWS
        N_ROWS = args.rows
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c5":
        return run_c5(args)
    if args.workload == "c3":
        return run_c3(args)
    if args.workload == "c4":
        return run_c4(args)
    if args.workload != "headline":
        return run_single_query_shapes(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
