#!/usr/bin/env python
"""Reads the [tm_trace] lines of a -DTM_TRACE run (tools/tokenmix_trace.sh) and prints, per role (warp) and per pair of
consecutive tags, how many clocks lie between them on average - where each role of CTA 0 spends a tile."""
import collections
import re
import sys

runs, cur = [], None
for line in open(sys.argv[1]):
    m = re.match(r"\[tm_trace\] mode (\d+) P (\d+) D (\d+)", line)
    if m:
        cur = {"hdr": line.strip(), "ev": collections.defaultdict(list)}
        runs.append(cur)
        continue
    m = re.match(r"\[tm_trace\] role (\d+) ev (\d+) tag (\d+) clk (\d+)", line)
    if m and cur is not None:
        cur["ev"][int(m.group(1))].append((int(m.group(3)), int(m.group(4))))
want = int(sys.argv[2]) if len(sys.argv) > 2 else len(runs) - 1
r = runs[want]
print(r["hdr"], f"(run {want} of {len(runs)})")
t0 = min(e[0][1] for e in r["ev"].values() if e)
t1 = max(e[-1][1] for e in r["ev"].values() if e)
print(f"kernel span {t1 - t0} clocks")
for role in sorted(r["ev"]):
    ev = r["ev"][role]
    gaps = collections.defaultdict(list)
    for (ta, ca), (tb, cb) in zip(ev, ev[1:]):
        gaps[(ta, tb)].append(cb - ca)
    desc = ", ".join(f"{a}->{b}: n={len(v)} avg={sum(v) / len(v):.0f}" for (a, b), v in sorted(gaps.items()))
    print(f"role {role:2d}: first {ev[0][1] - t0:6d} last {ev[-1][1] - t0:6d} | {desc}")
