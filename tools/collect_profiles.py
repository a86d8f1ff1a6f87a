#!/usr/bin/env python
"""Turns the raw output of tools/gpu_final.sh (gpurun_out/final/) into the tracked evidence under profiles/:
bench lines as they were printed, ncu launch lists (csv + a per-step table), and the full-capture summaries.
Usage: python tools/collect_profiles.py [round_tag]   (default r01)"""
import csv
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "final")       # r01; later rounds: gpurun_out/<tag>final (set in main)
DST = os.path.join(ROOT, "profiles")


def launches(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = {n: i for i, n in enumerate(rows[hi])}
    out = []
    for r in rows[hi + 1:]:
        if len(r) > h["Metric Value"]:
            out.append((r[h["Kernel Name"]], r[h["Grid Size"]], r[h["Block Size"]], float(r[h["Metric Value"]]) / 1e3))
    return out


def short(name):
    n = name.replace("void ", "").replace("cm::", "")
    return n.split("(")[0]


def step_table(ls, first_kernel):
    """rows of the LAST complete step: from the last launch of `first_kernel` back to the previous one."""
    idx = [i for i, l in enumerate(ls) if first_kernel in l[0]]
    if len(idx) < 2:
        return ls
    return ls[idx[-2]:idx[-1]]


def table(ls):
    tot = sum(l[3] for l in ls)
    lines = ["| kernel | grid | block | us | share |", "|---|---|---|---|---|"]
    for n, g, b, us in ls:
        lines.append(f"| `{short(n)}` | {g} | {b} | {us:.1f} | {100 * us / tot:.1f}% |")
    lines.append(f"| **sum** | | | **{tot:.1f}** | |")
    return "\n".join(lines)


def main():
    global SRC
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    if tag != "r01":
        SRC = os.path.join(ROOT, "gpurun_out", tag + "final")
    os.makedirs(DST, exist_ok=True)

    def cp(src, dst):
        p = os.path.join(SRC, src)
        if os.path.exists(p) and os.path.getsize(p) > 0:
            shutil.copyfile(p, os.path.join(DST, f"{tag}_{dst}"))
            return True
        print("missing:", src)
        return False

    for s, d in [("bench_n1.json", "bench_n1.json"), ("bench_reference.json", "bench_reference.json"),
                 ("bench_exact.json", "bench_exact_path.json"), ("bench_c1.json", "bench_c1_10Kx128_b1.json"),
                 ("bench_b1.json", "bench_exact_scan_b1_b8.json"), ("bench_shard_12.5M_l2.json", "bench_shard_12.5Mx768_l2.json"),
                 ("launches_tensor_path.csv", "launches_tensor_path.csv"), ("launches_exact_scan.csv", "launches_exact_scan.csv"),
                 ("idx_1m.json", "bench_indexes_1Mx768.json"), ("c3.json", "bench_c3_ivfpq_10Mx768.json"),
                 ("c4.json", "bench_c4_hnsw_1Mx768.json"), ("pytest.log", "pytest_gpu.log"),
                 ("summary_tensor.md", "ncu_full_tensor_path.md"), ("summary_scan.md", "ncu_full_exact_scan.md"),
                 ("summary_adc.md", "ncu_full_adc_scan.md"),
                 # round 2
                 ("bench_n1_l2.json", "bench_n1_l2.json"), ("bench_c3.json", "bench_c3.json"), ("bench_c4.json", "bench_c4.json"),
                 ("bench_c4_b8192.json", "bench_c4_b8192.json"), ("bench_c5_one_shard.json", "bench_c5_one_shard_12.5Mx768_l2.json"),
                 ("launches_tensor_path_l2.csv", "launches_tensor_path_l2.csv"), ("memcheck.log", "memcheck_all_paths.log"),
                 ("racecheck.log", "racecheck_all_paths.log"), ("gpu.txt", "gpu.txt")]:
        cp(s, d)
    # per-step launch tables
    md = []
    p = os.path.join(SRC, "launches_tensor_path.csv")
    if os.path.exists(p):
        ls = launches(p)
        md.append("## Tensor path: the launches of one step (512 queries x 1M x 768, K=100)\n")
        md.append("`ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 2 --warmup 3 "
                  "--no-cpu-baseline`; serialised, cold-cache durations: compare shares, not absolutes.\n")
        first = "prep_queries" if any("prep_queries" in l[0] for l in ls) else "preprocess_rows"
        md.append(table(step_table(ls, first)))
    p = os.path.join(SRC, "launches_exact_scan.csv")
    if os.path.exists(p):
        ls = [l for l in launches(p) if "flat_scan" in l[0]]
        if ls:
            us = sorted(l[3] for l in ls)
            md.append("\n## Exact path: `flat_scan_kernel` launches (8 queries per pass over 1M x 768)\n")
            md.append(f"{len(ls)} launches, median {us[len(us) // 2]:.1f} us, min {us[0]:.1f} us, max {us[-1]:.1f} us "
                      f"(grid {ls[0][1]}, block {ls[0][2]}).")
    with open(os.path.join(DST, f"{tag}_launch_tables.md"), "w") as f:
        f.write(f"# {tag} -- ncu launch lists, tabulated (raw csv beside this file)\n\n" + "\n".join(md) + "\n")
    # DRAM traffic of the dominant kernels from the full captures -> bench.py's roofline.traffic
    traffic = {"_source": f"ncu --set full captures of {tag} (profiles/{tag}_ncu_full_*.md): dram__bytes_read.sum + dram__bytes_write.sum per launch"}

    def dram(summary, kernel):
        p = os.path.join(SRC, summary)
        if not os.path.exists(p):
            return []
        vals, cur, rd = [], None, None
        for line in open(p):
            if line.startswith("## "):
                cur = line
            elif cur and kernel in cur and "dram__bytes_read.sum" in line:
                v, u = line.split("|")[2].strip(), line.split("|")[3].strip()
                rd = float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
            elif cur and kernel in cur and "dram__bytes_write.sum" in line and rd is not None:
                v, u = line.split("|")[2].strip(), line.split("|")[3].strip()
                vals.append(rd + float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u])
                rd = None
        return vals
    g = dram("summary_tensor.md", "flat_gemm_ts_kernel") or dram("summary_tensor.md", "flat_gemm_kernel")
    if g:
        traffic["flat_gemm_kernel_per_step_bytes"] = int(sum(g[-3:]))
        traffic["flat_gemm_kernel_launches"] = [int(x) for x in g]
    sc = dram("summary_scan.md", "flat_scan_kernel")
    if sc:
        traffic["flat_scan_kernel_per_launch_bytes"] = int(sc[-1])
    with open(os.path.join(DST, f"{tag}_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=2)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
