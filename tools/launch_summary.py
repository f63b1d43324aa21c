#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid, block): launches, total and mean time."""
import csv
import collections
import re
import sys

rows = collections.defaultdict(lambda: [0, 0.0])
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"^void |\(anonymous namespace\)::", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = val / 1e3 if unit in ("ns", "nsecond") else val if unit in ("us", "usecond") else val * 1e3
    k = (name[:70], r["Grid Size"], r["Block Size"])
    rows[k][0] += 1
    rows[k][1] += us
tot = sum(v[1] for v in rows.values())
print(f"total {tot / 1e3:.2f} ms over {sum(v[0] for v in rows.values())} launches")
for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{v[1] / 1e3:8.2f} ms {100 * v[1] / tot:5.1f}%  n={v[0]:5d}  mean {v[1] / v[0]:7.1f} us  grid {k[1]:>16} block {k[2]:>14}  {k[0]}")
