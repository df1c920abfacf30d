#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv` export: stall-reason totals and the hottest SASS lines.
Usage: scripts/ncu_source_stalls.py file_source.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = rows[1]
col = {c: i for i, c in enumerate(h)}
stall = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
data = []
for r in rows[2:]:
    if len(r) < len(h) - 1:
        continue
    try:
        data.append((int(r[col["# Samples"]]), int(r[col["Instructions Executed"]]), r))
    except ValueError:
        pass
tot = sum(d[0] for d in data)
inst = sum(d[1] for d in data)
print(f"kernel: {rows[0][1][:100]}")
print(f"samples {tot}, warp instructions executed {inst}, SASS lines {len(data)}")
agg = {c: 0 for c in stall}
for n, _, r in data:
    for c in stall:
        try:
            agg[c] += int(r[col[c]])
        except (ValueError, IndexError):
            pass
s = sum(agg.values())
print("stall reasons (all samples):")
for c, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print(f"  {c:24s} {100 * v / max(1, s):5.1f}%")
print("hottest SASS lines (samples, share, instr executed, text):")
for n, ie, r in sorted(data, key=lambda x: -x[0])[:top]:
    print(f"  {n:6d} {100 * n / max(1, tot):5.1f}% {ie:10d}  {r[col['Source']].strip()[:90]}")
