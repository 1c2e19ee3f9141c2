#!/usr/bin/env python
"""Top CUDA source lines of an ncu report by stall samples / executed instructions.
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > f.csv ; python tools/src_hot.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cur_file, agg, text = None, {}, {}
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        iS, iE = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    line = r[0]
    if line.strip().isdigit():
        cur_line = (cur_file, int(line))
        text[cur_line] = r[1]
    if r[2]:  # a SASS row (has an address)
        s = int(r[iS]) if r[iS].strip().isdigit() else 0
        e = int(r[iE]) if r[iE].strip().isdigit() else 0
        a = agg.setdefault(cur_line, [0, 0])
        a[0] += s
        a[1] += e
ts = sum(v[0] for v in agg.values()) or 1
te = sum(v[1] for v in agg.values()) or 1
print(f"samples {ts}  warp-instr {te}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * v[0] / ts:5.1f}% smp {100 * v[1] / te:5.1f}% ins  {k[0]}:{k[1]:4d}  {text.get(k, '')[:110].strip()}")
