"""Per-substep event clocks of CTA (0, 0) of the column-batched attention kernel (PB_ATTN_TRACE=1): where the pipeline waits.
    PB_ATTN_TRACE=1 python scripts/trace_attn.py [case]
columns (clocks relative to the first event): pcld = per-column TMA issued, shld = shared-stage TMA issued, w1: pc_full seen /
s_free seen / S issued, w10: pc_full seen (P.C2) / t_full seen (T.C1), compute: start / s_full seen / math done / t_empty seen / t_full arrived"""
import ctypes as C, os, sys, runpy, torch
os.environ.setdefault("PB_ATTN_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
case = sys.argv[1] if len(sys.argv) > 1 else "jvp"
sys.argv = ["bench_attn.py", "--shapes", "one:" + case, "--reps", "1"]
ns = runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bench_attn.py"))
from diffusion_pullback_b200 import _native as N
buf = (C.c_longlong * (512 * 16))()
n = N.raw().pbk_attn16_trace_read(buf, 512 * 16)
ev = [[buf[u * 16 + e] for e in range(16)] for u in range(512)]
t0 = min(v for r in ev for v in r if v)
names = ["w1_pc", "w1_sfree", "w1_iss", "w10_pc", "w10_tfull", "c_sfull", "c_math", "c_tempty", "c_tfull", "c_start", "pcld", "shld", "w1_el", "w1_mma", "w1_cmt", "x"]
order = [10, 11, 0, 1, 12, 13, 14, 2, 9, 5, 6, 7, 8, 3, 4]
print("u   " + " ".join(f"{names[e]:>9}" for e in order))
for u in range(int(os.environ.get("TRACE_FROM", "100")), int(os.environ.get("TRACE_TO", "140"))):
    print(f"{u:<3} " + " ".join(f"{(ev[u][e] - t0) if ev[u][e] else -1:>9}" for e in order))
