#!/usr/bin/env python
"""Hot-region breakdown of an ncu source page (SASS) CSV:  ncu -i X.ncu-rep --page source --csv > f.csv
    python tools/sass_hot.py f.csv [chunks]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_idx = [i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r]
h = rows[hdr_idx[0]]
end = hdr_idx[1] - 1 if len(hdr_idx) > 1 else len(rows)
data = [r for r in rows[hdr_idx[0] + 1:end] if len(r) == len(h)]
iS, iSm, iE = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
num = lambda x: int(x) if x.strip().isdigit() else 0
tot_s = sum(num(r[iSm]) for r in data) or 1
tot_e = sum(num(r[iE]) for r in data) or 1
print("samples", tot_s, "warp-instr", tot_e, "sass lines", len(data))
n = len(data)
ch = max(1, n // int(sys.argv[2] if len(sys.argv) > 2 else 36))
for c in range(0, n, ch):
    seg = data[c:c + ch]
    s = sum(num(r[iSm]) for r in seg)
    e = sum(num(r[iE]) for r in seg)
    ops = collections.Counter((r[iS].split()[1] if r[iS].startswith("@") else r[iS].split()[0]) for r in seg if r[iS].split())
    print(f"{c:5d} samples {100 * s / tot_s:5.1f}% instr {100 * e / tot_e:5.1f}%", ops.most_common(6))
