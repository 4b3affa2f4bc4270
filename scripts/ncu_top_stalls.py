"""Top stall-sample SASS instructions of one kernel of an ncu report:
  ncu -i <rep> --page source --csv --kernel-id :::<n> > src.csv ; python scripts/ncu_top_stalls.py src.csv [ntop]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if "Address" in r)
hdr = rows[h]
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[0] != "Address"]
isrc, iss, ie = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
num = lambda s: int(float(s)) if s else 0
tot = sum(num(d[iss]) for d in data)
print(rows[0][1][:120] if len(rows[0]) > 1 else "", "| total samples", tot, "| instructions", len(data))
top = sorted(range(len(data)), key=lambda k: -num(data[k][iss]))[:ntop]
for k in sorted(top):
    d = data[k]
    print(f"{k:5d} {100.0 * num(d[iss]) / max(tot, 1):5.1f}% x{num(d[ie]):7d}  {d[isrc][:110]}")
