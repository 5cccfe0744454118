#!/usr/bin/env python
"""Top stall-sampled SASS instructions from `ncu -i X.ncu-rep --page source --csv` output.
usage: ncu_hot.py src.csv [section_index] [topN]"""
import csv, sys
path = sys.argv[1]; sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(path)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
secs.append(len(rows))
s, e = secs[sec], secs[sec + 1]
print(rows[s][1])
hdr = rows[s + 1]
body = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
ci = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
order = sorted(range(len(body)), key=lambda i: -int(body[i][ci["# Samples"]] or 0))[:top]
for i in order:
    r = body[i]
    n = int(r[ci["# Samples"]] or 0)
    st = sorted(((int(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {100.0*n/tot:5.1f}% {r[ci['Source']].strip()[:90]:90s} exec={r[ci['Instructions Executed']]:>9s} " + " ".join(f"{c[6:]}={v}" for v, c in st if v))
