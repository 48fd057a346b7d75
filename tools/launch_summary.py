#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols = rows[hdr]
data = rows[hdr + 1:]
ki, vi, ui = cols.index("Kernel Name"), cols.index("Metric Value"), cols.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in data:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).split("::")[-1]
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    if r[ui] in ("ns", "nsecond"):
        v /= 1e3
    elif r[ui] in ("ms", "msecond"):
        v *= 1e3
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:64]:64} n={v[0]:4d} us={v[1]:10.1f} share={v[1] / tot * 100:5.1f}%")
