#!/usr/bin/env python3
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, mean duration and share of the total per kernel.
  python tools/launch_summary.py gpurun_out/launches_r2c_bench.csv "<command the list was taken from>" """
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
cmd = sys.argv[2] if len(sys.argv) > 2 else "python bench.py --steps 3 --warmup 3"
lines = open(path, errors="replace").read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(lines[start:]))
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    name = r["Kernel Name"]
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"launch list of `{cmd}` under ncu (gpu__time_duration.sum, ns; cold cache, serialised: compare shares)")
print(f"total {total/1e3:.1f} us over {sum(cnt.values())} launches")
for name in sorted(tot, key=lambda k: -tot[k]):
    print(f"{100*tot[name]/total:6.2f} % {cnt[name]:6d} x {tot[name]/cnt[name]/1e3:10.2f} us  {name[:100]}")
