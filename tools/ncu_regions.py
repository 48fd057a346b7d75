#!/usr/bin/env python
"""Per-role view of an .ncu-rep of gemm_tc_kernel: raw headline metrics, then warp-stall samples grouped by the SASS
region they fall in (regions are told apart by marker instructions), and the top stall PCs of each region."""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[-1]))
print("kernel:", d.get("Kernel Name", "?")[:110])
for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__registers_per_thread",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active"]:
    print(f"  {k} = {d.get(k, 'n/a')}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ia, isrc, iall, iex = h.index("Address"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > iex and r[ia].startswith("0x")]
data = data[:len(data) // 2] if len(data) > 2 and data[0][ia] == data[len(data) // 2][ia] else data
tot = sum(int(r[iall] or 0) for r in data)
print(f"  {len(data)} SASS instructions, {tot} samples")
top = sorted(range(len(data)), key=lambda i: -int(data[i][iall] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for i in top:
    r = data[i]
    ctx = ""
    for j in range(i, max(i - 40, 0), -1):   # nearest marker before the PC
        m = re.search(r"(UTCHMMA|UTMALDG|UTMASTG|UTCBAR|LDTM|UTMACMDFLUSH|UCGABAR_WAIT|MUFU\.\w+|STS\.128|LDS\.128|SYNCS\.ARRIVE\S*)", data[j][isrc])
        if m:
            ctx = f"[after {m.group(1)} @{j}]"
            break
    print(f"    {int(r[iall] or 0):5d} {100 * int(r[iall] or 0) / max(tot, 1):5.1f}%  idx {i:5d} ex={r[iex]:>8}  {r[isrc].strip()[:70]:70s} {ctx}")
