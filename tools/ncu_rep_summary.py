#!/usr/bin/env python
"""Summarise one .ncu-rep holding several kernels (tools/prof_kernels.py under ncu --set full) into markdown and
update profiles/ncu_traffic.json (DRAM read + write bytes AND L2 bytes per launch).
    python tools/ncu_rep_summary.py gpurun_out/prof_r02a.ncu-rep r02a
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from ncu_summary import KEYS  # noqa: E402

rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
ki = h.index("Kernel Name")
seen = {}
for r in data:  # keep the LAST launch of every kernel name (the one after the warm-up and the L2 flush)
    seen[r[ki].split("(")[0].replace("void ", "")] = r
out = [f"# ncu --set full — {tag}", "",
       "`ncu --set full --clock-control none --import-source on -k regex:tpg -c 60 python tools/prof_kernels.py`: one launch of every "
       "hot kernel at a BASELINE configs[1] shape (second launch, L2 flushed before it; K2 on real generator activations).", ""]
# algorithmic bytes (SURVEY.md §8d formulas) of the shapes tools/prof_kernels.py launches
B_ = 8
ALG = {"group_bwd_staged_kernel": 4 * B_ * (64 * 32768 + 32768 + 64 * 2048), "group_fwd_kernel": 4 * B_ * (64 * 2048 + 32768 + 64 * 32768),
       "group_reduce_fwd_smem_kernel": 4 * B_ * (64 * 2048 + 32768 + 2 * 64 * 2048), "knn_feat_tc_kernel": 4 * B_ * 64 * 4096 + 12 * B_ * 2048 * 12,
       "fps_reg_kernel": 12 * B_ * 2048 + 4 * B_ * 512, "ball_query_kernel": 12 * B_ * (8192 + 1024) + 4 * B_ * 1024 * 32,
       "grid_knn_kernel": 4 * B_ * 3 * 2 * 8192 + 12 * B_ * 8192 * 16, "grid_nn1_kernel": 20 * B_ * 2 * 8192,
       "csr_cluster_kernel": 4 * B_ * (2 * 32768 + 2048),
       "group_assemble_kernel": 4 * B_ * (3 * 256 + 3 * 256 + 2 * 256 * 256 + 256 * 32 + (3 + 512) * 256 * 32),
       "edge_affine_kernel": 4 * B_ * (3 * 16 * 2048 + 2048 * 20 + 16 * 2048 * 20)}
traffic = {"_source": f"{tag}: dram__bytes_read.sum + dram__bytes_write.sum and lts__t_bytes.sum of one launch per kernel "
                      "(ncu --set full --clock-control none, L2 flushed before the launch; tools/prof_kernels.py)"}
for name, r in seen.items():
    out += [f"## `{name[:100]}`", "", "| metric | unit | value |", "|---|---|---:|"]
    for k in KEYS:
        if k in h:
            out.append(f"| {k} | {units[h.index(k)]} | {r[h.index(k)]} |")
    out.append("")

    def val(k):
        if k not in h:
            return None
        v, u = float(r[h.index(k)].replace(",", "")), units[h.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    rd, wr, l2 = val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), val("lts__t_bytes.sum")
    import re

    short = re.sub(r"<[^<>]*>$", "", name.replace("<unnamed>::", "").replace("unnamed>::", "").replace("tpg::", "")).strip()
    traffic[short] = {"dram_bytes_per_launch": (rd or 0) + (wr or 0), "dram_read": rd, "dram_write": wr, "l2_bytes_per_launch": l2,
                      "duration_us": (val("gpu__time_duration.sum") or 0) / 1e3 if units[h.index("gpu__time_duration.sum")] == "ns" else val("gpu__time_duration.sum"),
                      "launch": name[:120], "alg_bytes_of_this_launch": ALG.get(short)}
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_full.md"), "w").write("\n".join(out) + "\n")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print("kernels:", list(seen))
