"""Per-shape table of a PB_PROFILE_DUMP file (one line per probed contraction launch: us, GF, shape):
    PB_PROFILE_DUMP=gpurun_out/probe.txt python bench.py --no-cpu-baseline --no-ref-gpu --no-secondary --steps 2 --warmup 3
    python scripts/per_shape_table.py gpurun_out/probe.txt"""
import sys
from collections import defaultdict

agg = defaultdict(lambda: [0, 0.0, 0.0])
for line in open(sys.argv[1]):
    us, _, gf, _, shape = line.rstrip("\n").split(" ", 4)
    a = agg[shape]
    a[0] += 1; a[1] += float(us); a[2] += float(gf)
tot = sum(a[1] for a in agg.values())
print(f"{'shape':104s} {'n':>3s} {'total us':>10s} {'share':>6s} {'avg us':>8s} {'TF/s':>7s}")
for shape, (n, us, gf) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{shape:104s} {n:3d} {us:10.1f} {100 * us / tot:5.1f}% {us / n:8.1f} {gf / us * 1e3:7.1f}")
