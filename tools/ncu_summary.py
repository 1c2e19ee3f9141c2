#!/usr/bin/env python
"""Summarise gpurun_out ncu artefacts into profiles/<tag>_*.md (run in the build container).
    python tools/ncu_summary.py r01a
"""
import collections
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(tag):
    path = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if v != v:  # ncu occasionally reports nan for a launch
            continue
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        k = row["Kernel Name"].split("(")[0].replace("void ", "")
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"# ncu launch list — {tag}", "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 1 --warmup 1 "
           "--no-e2e --no-cpu-baseline` (2 passes of the fluid step schedule, batch 8). Per-launch times are cold-cache "
           "and serialised: compare SHARES with bench.py's per-op table, not absolutes.", "",
           f"total {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches", "",
           "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k[:90]}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |")
    return "\n".join(out) + "\n"


def full(tag):
    out = [f"# ncu --set full summaries — {tag}", "",
           "`ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 2` on the same bench command; "
           "values per launch.", ""]
    for rep in sorted(glob.glob(os.path.join(OUT, f"prof_*_{tag}.ncu-rep"))):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        if len(rows) < 3:
            continue
        h, units, data = rows[0], rows[1], rows[2:]
        name = os.path.basename(rep)[5:-len(f"_{tag}.ncu-rep")]
        ki = h.index("Kernel Name")
        out += [f"## {name}", "", "kernel: `" + data[0][ki][:120] + "`", "", "| metric | unit | " + " | ".join(
            f"launch {i}" for i in range(len(data))) + " |", "|---|---|" + "---:|" * len(data)]
        for k in KEYS:
            if k in h:
                i = h.index(k)
                out.append(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
        out.append("")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    tag = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    s = launches(tag)
    if s:
        open(os.path.join(PROF, f"{tag}_launches.md"), "w").write(s)
    open(os.path.join(PROF, f"{tag}_ncu_full.md"), "w").write(full(tag))
    for f in (f"bench_{tag}.json", f"bench_ref_{tag}.json", f"bench_{tag}.err"):
        p = os.path.join(OUT, f)
        if os.path.exists(p):
            open(os.path.join(PROF, f), "w").write(open(p).read())
    print("wrote profiles for", tag)
