#!/usr/bin/env python
"""ncu_source_hot.py <source-page.csv> [top] — from `ncu -i rep --page source --csv`: stall reasons of the whole kernel, the SASS instructions with the
most warp-state samples, and samples by opcode relative to how often the opcode executes (anomalies: a few instructions that hold many warps)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print(rows[0][1][:100])
hdr = rows[1]; data = rows[2:]
ia, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
ci = [hdr.index(c) for c in cols]
tot = sum(int(r[isamp] or 0) for r in data); totex = sum(int(r[iex] or 0) for r in data)
print("samples", tot, "warp instructions", totex, "static instructions", len(data))
print("by reason %:", {c[6:]: round(100 * sum(int(r[k] or 0) for r in data) / tot, 1) for c, k in zip(cols, ci) if sum(int(r[k] or 0) for r in data) * 200 > tot})
byop = collections.defaultdict(lambda: [0, 0])
for r in data:
    m = re.match(r"\s*(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[ia]); op = m.group(1) if m else "?"
    byop[op][0] += int(r[isamp] or 0); byop[op][1] += int(r[iex] or 0)
print(f"{'opcode':24s} {'samples%':>8s} {'exec%':>7s} {'ratio':>6s}")
for op, v in sorted(byop.items(), key=lambda kv: -kv[1][0])[:18]:
    print(f"{op:24s} {100 * v[0] / tot:8.1f} {100 * v[1] / totex:7.1f} {(v[0] / tot) / (v[1] / totex) if v[1] else 0:6.2f}")
print("hottest instructions:")
for i, r in sorted(enumerate(data), key=lambda kv: -int(kv[1][isamp] or 0))[:top]:
    rs = {c[6:]: int(r[k] or 0) for c, k in zip(cols, ci) if int(r[k] or 0)}
    t3 = sorted(rs.items(), key=lambda kv: -kv[1])[:3]
    print(f"  #{i:5d} {100 * int(r[isamp] or 0) / tot:5.2f}%  x{int(r[iex] or 0):>9d}  {r[ia].strip()[:64]:64s} {t3}")
