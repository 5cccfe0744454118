#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_summary.py launches.csv [steps]   (steps divides the totals -> ms per step)"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h = rows[hdr]; ci = {k: i for i, k in enumerate(h)}
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) < len(h): continue
    name = re.sub(r'\(.*', '', r[ci['Kernel Name']])[:80]
    v = float(r[ci['Metric Value']].replace(',', '')); u = r[ci['Metric Unit']]
    v = v / 1000 if u == 'us' else (v / 1e6 if u == 'ns' else v)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"| kernel | launches | ms{' / step' if steps != 1 else ''} | share |\n|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1] / steps:.3f} | {100 * a[1] / tot:.1f} % |")
print(f"| total | | {tot / steps:.3f} | |")
