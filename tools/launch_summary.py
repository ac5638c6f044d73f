#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (second half = warm run)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
skip_first_half = len(sys.argv) < 3 or sys.argv[2] != "all"
lines = [l for l in open(path) if not l.startswith("==")]
allr = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"us": 1e3, "ms": 1e6, "ns": 1.0, "s": 1e9}.get(row["Metric Unit"], 1.0)
    allr.append((row["Kernel Name"], v))
sel = allr[len(allr) // 2:] if skip_first_half else allr
tot, cnt = collections.defaultdict(float), collections.Counter()
for name, v in sel:
    short = re.sub(r"\(.*", "", name)[:64]
    tot[short] += v
    cnt[short] += 1
T = sum(tot.values())
print(f"launches {len(sel)}  total {T/1e6:.3f} ms")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{v/1e6:9.3f} ms {100*v/T:5.1f}%  x{cnt[k]:4d}  avg {v/cnt[k]/1e3:8.1f} us  {k}")
