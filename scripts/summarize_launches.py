"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0       # launches to skip (e.g. weight packing + primal)
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
    rows.append((re.sub(r"\(.*", "", r["Kernel Name"]), ns))
rows = rows[skip:]
tot = sum(ns for _, ns in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
print(f"launches {len(rows)}  total {tot/1e6:.3f} ms")
print(f"{'kernel':60s} {'count':>7s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:60]:60s} {c:7d} {ns/1e6:10.3f} {100*ns/tot:6.1f}% {ns/c/1e3:9.1f}")
