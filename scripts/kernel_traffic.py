"""DRAM traffic per launch of the contraction kernels over ONE subspace iteration, from an ncu metrics pass:

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/traffic.csv python scripts/profile_iter.py --workload sd15_mid_k5_i50 --iters 2
    python scripts/kernel_traffic.py gpurun_out/traffic.csv sd15_mid_k5_i50 "<source note>"      -> profiles/kernel_traffic.json

The last iteration (everything after the previous rotate_k launch) is summarised; bench.py reads the json for
`roofline.traffic`."""
import csv
import json
import os
import re
import sys
from collections import defaultdict

path, workload, source = sys.argv[1], sys.argv[2], sys.argv[3]
lines = [l for l in open(path, newline="") if not l.startswith("==")]
launch = {}
order = []
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
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2      # iterations in the pass; with P problem slots P rotate_k launches close each
per = max(1, len(ends) // iters)
rows = rows[ends[-per - 1] + 1: ends[-1] + 1] if len(ends) > per else rows
agg = defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    key = "gemm_tc_kernel" if "gemm_tc_kernel" in r["name"] else "attn_lin_kernel" if ("attn_lin_kernel" in r["name"] or "attn16_kernel" in r["name"]) else "other"
    a = agg[key]
    a[0] += 1
    a[1] += r.get("dram__bytes_read.sum", 0.0) + r.get("dram__bytes_write.sum", 0.0)
    a[2] += r.get("gpu__time_duration.sum", 0.0)
out_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "kernel_traffic.json")
data = json.load(open(out_path)) if os.path.exists(out_path) else {}
data[workload] = {k: {"launches_per_iter": n, "dram_bytes_per_launch": b / n, "dram_bytes_per_iter": b, "us_per_iter_under_ncu": t,
                      "source": source} for k, (n, b, t) in agg.items()}
json.dump(data, open(out_path, "w"), indent=1)
print(json.dumps(data[workload], indent=1))
