#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time and share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
nsteps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = []
with open(path) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"<.*$", "", name) if name.startswith("at::") or "at::native" in name else name
    rows.append((name[:90], v * scale))
agg = defaultdict(lambda: [0, 0.0])
for n, ms in rows:
    agg[n][0] += 1
    agg[n][1] += ms
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {len(rows)} launches, {tot:.1f} ms total device time ({nsteps:g} steps -> {tot / nsteps:.1f} ms/step, cold-cache serialised)")
print(f"# {'kernel':90s} {'launches/step':>13s} {'ms/step':>9s} {'share':>6s}")
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:92s} {c / nsteps:13.1f} {ms / nsteps:9.3f} {100 * ms / tot:5.1f}%")
