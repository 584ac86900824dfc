#!/usr/bin/env python3
"""Per-CUDA-source-line stall samples / instruction counts from `ncu -i rep --page source --csv --print-source sass,cuda`.
usage: ncu_cuda_lines.py file.csv [top] [n_pods]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
npods = float(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(path, errors="replace")))
cur_file, hdr = "?", None
agg = {}
stall_names = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    a = agg.setdefault((cur_file, ln), [0, 0, r[1].strip(), defaultdict(int)])
    a[0] += int(r[si] or 0)
    a[1] += int(r[ii] or 0)
    for c in stall_cols:
        v = int(r[c] or 0)
        if v:
            a[3][hdr[c][6:]] += v
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print(f"total samples {tot}, warp instructions {toti}" + (f", {toti / npods:.0f} / pod" if npods else ""))
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    why = ",".join(f"{k}:{v}" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
    per = f" {a[1] / npods:6.1f}/pod" if npods else ""
    print(f"{100 * a[0] / tot:5.1f}% smp {100 * a[1] / toti:5.1f}% ins{per}  {f}:{ln:<5} {a[2][:70]:70s} {why}")
