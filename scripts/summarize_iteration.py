"""Per-kernel summary (count / time / share / DRAM bytes) of the LAST subspace iteration in an ncu metrics pass:
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/x.csv python scripts/profile_iter.py --iters 2 --slots 5
    python scripts/summarize_iteration.py gpurun_out/x.csv [iterations in the pass]"""
import csv
import re
import sys
from collections import defaultdict

lines = [l for l in open(sys.argv[1], newline="") if not l.startswith("==")]
launch, order = {}, []
for r in csv.DictReader(lines):
    i = int(r["ID"])
    if i not in launch:
        launch[i] = {"name": re.sub(r"\(.*", "", r["Kernel Name"])}
        order.append(i)
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    launch[i][r["Metric Name"]] = v * scale
rows = [launch[i] for i in order]
ends = [j for j, r in enumerate(rows) if "rotate_k" in r["name"]]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2       # iterations in the pass (P rotate_k launches close each one)
per = len(ends) // iters
rows = rows[ends[-per - 1] + 1: ends[-1] + 1] if len(ends) > per else rows
agg = defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    a = agg[r["name"]]
    a[0] += 1
    a[1] += r.get("gpu__time_duration.sum", 0.0)
    a[2] += r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
print(f"launches {len(rows)}  total {tot / 1e3:.3f} ms   DRAM {sum(a[2] for a in agg.values()) / 1e9:.2f} GB")
print(f"{'kernel':66s} {'count':>5s} {'total ms':>9s} {'share':>6s} {'avg us':>8s} {'DRAM MB/launch':>15s} {'DRAM GB/s':>10s}")
for n, (c, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:66]:66s} {c:5d} {us / 1e3:9.3f} {100 * us / tot:5.1f}% {us / c:8.1f} {by / c / 1e6:15.1f} {by / us / 1e3:10.0f}")
