#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list
by kernel: launches, device time, share of the step and DRAM bytes."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols = rows[hdr]
ki, mi, vi, ui, ii = (cols.index(c) for c in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
agg = collections.defaultdict(lambda: {"n": set(), "us": 0.0, "rd": 0.0, "wr": 0.0})
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).split("::")[-1]
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    u = r[ui]
    a = agg[name]
    a["n"].add(r[ii])
    if r[mi].startswith("gpu__time_duration"):
        a["us"] += v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
    elif r[mi].startswith("dram__bytes"):
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
        a["rd" if "read" in r[mi] else "wr"] += v * scale
tot = sum(a["us"] for a in agg.values())
nl = sum(len(a["n"]) for a in agg.values())
print(f"ncu launch list: total {tot:.0f} us over {nl} launches (cold-cache, serialised: compare shares)")
print(f"{'kernel':46} {'n':>5} {'us':>9} {'share':>6} {'dram_rd_MB':>10} {'dram_wr_MB':>10}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    print(f"{k[:46]:46} {len(a['n']):5d} {a['us']:9.1f} {a['us'] / tot * 100:5.1f}% {a['rd']:10.1f} {a['wr']:10.1f}")
if len(sys.argv) > 2:      # launch_summary.py launches.csv out.json [commit]: the traffic file bench.py stamps roofline.traffic with
    import json
    fam = {f: [a for k, a in agg.items() if k.startswith(f)] for f in ("gemm_tc_kernel", "token_mix_kernel")}
    out = {"kernel": "gemm_tc_kernel (all launches of one B/32 batch-256 training step, fused token-mixing schedule)",
           "dram_bytes_per_step": sum(a["rd"] + a["wr"] for a in fam["gemm_tc_kernel"]) * 1e6,
           "launches": sum(len(a["n"]) for a in fam["gemm_tc_kernel"]),
           "share_of_step": sum(a["us"] for a in fam["gemm_tc_kernel"]) / tot,
           "source": f"profiles/{sys.argv[1].split('/')[-1]} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum)",
           "commit": sys.argv[3] if len(sys.argv) > 3 else None,
           "token_mix_kernel": {"dram_bytes_per_step": sum(a["rd"] + a["wr"] for a in fam["token_mix_kernel"]) * 1e6,
                                "launches": sum(len(a["n"]) for a in fam["token_mix_kernel"]),
                                "share_of_step": sum(a["us"] for a in fam["token_mix_kernel"]) / tot}}
    json.dump(out, open(sys.argv[2], "w"), indent=1)
for fam in ("gemm_tc_kernel", "token_mix_kernel"):
    f = [a for k, a in agg.items() if k.startswith(fam)]
    if f:
        us = sum(a["us"] for a in f)
        print(f"{fam} (all instantiations): {us:.0f} us = {us / tot * 100:.1f}% of the step, DRAM traffic "
              f"{sum(a['rd'] + a['wr'] for a in f) / 1e3:.3f} GB per step")
