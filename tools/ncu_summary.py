#!/usr/bin/env python
"""Summaries of the ncu artefacts that gpurun brings back, for profiles/ (run HERE, no GPU needed).

    python tools/ncu_summary.py launches gpurun_out/r01c_launches.csv  > profiles/r01c_launch_list_summary.txt
    python tools/ncu_summary.py full gpurun_out/r01c_top.ncu-rep       > profiles/r01c_ncu_top_kernels.json

`launches`: per-kernel launch counts / total time / share of all captured device time (the csv of
`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).
`full`: the metrics DESIGN.md quotes, per captured launch (`ncu -i rep --page raw --csv`).
"""
import collections
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "launch__shared_mem_per_block_dynamic", "launch__block_size", "launch__grid_size", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_sleeping",
        "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_branch_resolving"]


def short(name):
    name = name.replace("void ", "").replace("fesr::", "")
    return name.split("(")[0]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)      # -> us
        k = short(row["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{len(agg)} kernels, {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms of device time (cold-cache, serialised)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {k[:78]:80s} n={v[0]:5d}  {v[1]:11.1f} us  {v[1] / v[0]:9.2f} us/launch  {100 * v[1] / tot:5.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        ent = {"Kernel Name": r[hdr.index("Kernel Name")]}
        for m in KEEP:
            if m in hdr:
                i = hdr.index(m)
                ent[m] = f"{r[i]} {units[i]}".strip()
        res.append(ent)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
