#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (last step only)."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    rows.append((r["Kernel Name"], us, r.get("Grid Size", ""), r.get("Block Size", "")))
n = len(rows) // steps
last = rows[-n:]
agg = defaultdict(lambda: [0, 0.0])
for name, us, g, b in last:
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    short = re.sub(r"\(anonymous namespace\)::", "", short)
    agg[short][0] += 1
    agg[short][1] += us
tot = sum(v[1] for v in agg.values())
print(f"launches in last step: {n}, total device time {tot / 1e3:.2f} ms (serialised, cold cache)")
print(f"{'kernel':70s} {'n':>5s} {'ms':>9s} {'share':>7s}")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {c:5d} {us / 1e3:9.3f} {100 * us / tot:6.1f}%")
