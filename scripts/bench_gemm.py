"""Per-shape timing of the tcgen05 GEMM on the shapes one pullback iteration launches (GPU box).
  PB_TRACE_GEMM=1 python scripts/profile_iter.py --iters 1 | grep ^GEMM > gpurun_out/gemm_shapes.txt   (shape log)
  python scripts/bench_gemm.py gpurun_out/gemm_shapes.txt
Each distinct shape is timed with CUDA events (20 launches, operands rotated through > 126 MB so L2 is cold-ish)."""
import collections
import ctypes as C
import re
import sys

import torch

sys.path.insert(0, ".")
from diffusion_pullback_b200 import _native as N

shapes = collections.Counter()
for line in open(sys.argv[1]):
    if line.startswith("GEMM"):
        shapes[tuple(int(v) for v in re.findall(r"=(-?\d+)", line))] += 1
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
f = N.leaf("pbk_gemm")
tune = N.raw().pb_gemm_tune
cfgs = [("default", (0, 1, 1)), ("nosplit", (0, 0, 1)), ("no160", (0, 1, 0)), ("bn128", (128, 1, 0))]
ws = torch.empty(8 << 20, device="cuda")

if len(sys.argv) > 2:
    cfgs = [c for c in cfgs if c[0] in sys.argv[2:]]
rows = []
for key, cnt in shapes.items():
    M, Nn, K0, K1, nseg, nb, nh, conv, H, W, res = key
    K = K0 + K1
    nbat = nb * nh if not conv else 1
    flops = 2.0 * M * Nn * (9 * K if conv else K) * nbat
    nrot = 4
    r4 = lambda v: (v + 3) // 4 * 4
    A = torch.randn(nrot, nbat, M, r4(K0), device="cuda")
    A2 = torch.randn(nrot, nbat, M, max(r4(K1), 4), device="cuda")
    B = torch.randn(nrot, nbat, Nn, (9 * K0 if conv else r4(K0)), device="cuda")
    B2 = torch.randn(nrot, nbat, Nn, max(r4(K1), 4), device="cuda")
    D = torch.empty(nrot, nbat, M, (Nn + 3) // 4 * 4, device="cuda")
    gs = []
    for r in range(nrot):
        g = N.PbGemm()
        g.M, g.N, g.nseg = M, Nn, nseg
        s = g.seg[0]
        s.A, s.lda, s.sAb, s.sAh, s.B, s.ldb, s.sBb, s.sBh, s.K = A[r].data_ptr(), A.shape[-1], M * A.shape[-1], 0, B[r].data_ptr(), B.shape[-1], Nn * B.shape[-1], 0, K0
        if nseg > 1:
            s = g.seg[1]
            s.A, s.lda, s.sAb, s.sAh, s.B, s.ldb, s.sBb, s.sBh, s.K = A2[r].data_ptr(), A2.shape[-1], M * A2.shape[-1], 0, B2[r].data_ptr(), B2.shape[-1], Nn * B2.shape[-1], 0, K1
        g.D, g.ldd, g.sDb = D[r].data_ptr(), D.shape[-1], M * D.shape[-1]
        if res:
            g.R, g.ldr, g.sRb, g.beta = D[r].data_ptr(), D.shape[-1], M * D.shape[-1], 1.0
        g.alpha, g.nb, g.nh, g.conv, g.H, g.W, g.round_tf32 = 1.0, (nb if conv else nbat), 1, conv, H, W, 1
        g.ws, g.ws_floats = ws.data_ptr(), ws.numel()
        gs.append(g)
    uss = []
    for name, cfg in cfgs:
        tune(*cfg)
        for g in gs:
            err = f(C.byref(g), st)
            assert err is None, (err, key)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            f(C.byref(gs[i % nrot]), st)
        e1.record()
        torch.cuda.synchronize()
        uss.append(e0.elapsed_time(e1) * 1e3 / 20)
    us = uss[0]
    byts = 4.0 * (M * K * nbat + Nn * (9 * K if conv else K) * nbat + M * Nn * nbat * (2 if res else 1))
    rows.append((us * cnt, cnt, key, uss, flops / 1e6, byts / 1e3))
tot = [sum(r[3][i] * r[1] for r in rows) for i in range(len(cfgs))]
print("configs:", [c[0] for c in cfgs], "totals ms:", [round(t / 1e3, 2) for t in tot], "launches", sum(r[1] for r in rows))
print("  share  count  us/launch per config | best TF/s  GB/s |  M N K0 K1 nseg nb nh conv H W res")
for t, cnt, key, uss, fl, by in sorted(rows, reverse=True):
    b = min(uss)
    print(f"{100*t/tot[0]:6.1f}% {cnt:5d}  " + " ".join(f"{u:8.1f}" for u in uss) + f" | {fl/b:7.1f} {by/b:7.0f} |  {key}")
