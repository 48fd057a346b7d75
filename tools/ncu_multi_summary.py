#!/usr/bin/env python
"""Summarise every launch of an .ncu-rep (`ncu --set full`): duration, tensor-pipe and DRAM utilisation, DRAM bytes,
issue statistics and the largest warp-stall reasons."""
import csv
import re
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
seen = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d.get("Kernel Name", "?")
    m = re.search(r"(\w+<[^>]*>)\(", name)
    short = m.group(1) if m else name.split("(")[0][-60:]
    seen[short] = seen.get(short, 0) + 1
    if len(sys.argv) > 2 and seen[short] != int(sys.argv[2]):
        continue      # keep the n-th launch of every kernel (warm caches)
    print(f"=== {short}  (launch {seen[short]} of this kernel in the capture)")
    for k in keys:
        if k in d:
            print(f"  {k} = {d[k]} {rows[1][hdr.index(k)]}")
    stalls = sorted(((float(v), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k, v in d.items()
                     if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v not in ("", "0")),
                    reverse=True)
    print("  stall samples:", ", ".join(f"{n}={int(c)}" for c, n in stalls[:8]))
