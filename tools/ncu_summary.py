#!/usr/bin/env python
"""Summarise an .ncu-rep (one launch): key raw metrics, saturated units and the top stall PCs."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[-1]))
print("kernel:", d.get("Kernel Name", "?")[:100])
keys = ["gpu__time_duration.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__registers_per_thread",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for k in keys:
    print(f"  {k} = {d.get(k, 'n/a')}")
stalls = sorted(((int(float(v)), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k, v in d.items()
                 if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued") and v not in ("", "0")), reverse=True)
print("  stall samples:", ", ".join(f"{n}={c}" for c, n in stalls[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc, iss, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [r for r in rows[2:] if len(r) > iex]
tot = sum(int(r[iss] or 0) for r in data)
print(f"  top stall PCs (of {tot} samples):")
for r in sorted(data, key=lambda r: -int(r[iss] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]:
    print(f"    {int(r[iss]):5d} {100 * int(r[iss]) / max(tot, 1):5.1f}%  ex={r[iex]:>8}  {r[isrc][:100]}")
