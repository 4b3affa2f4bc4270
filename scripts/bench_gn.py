"""Timing of the fp16-tangent elementwise kernels (GroupNorm / LayerNorm / GEGLU linearisations) on the SD-1.5 layer shapes,
CUDA events over back-to-back launches on rotating buffers (warm instruction cache, tensors larger than what one launch
leaves in L2 only for the big shapes): algorithmic bytes / time against the HBM peak.  `--nb` images (k, or slots * k)."""
import argparse, ctypes as C, json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_pullback_b200 import _native as N
ap = argparse.ArgumentParser(); ap.add_argument("--nb", type=int, default=5); a = ap.parse_args()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
nfl = N.raw().pbk_gn_tmp_floats
nfl.restype = C.c_size_t
nb, res = a.nb, []


def timeit(fn, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps           # us


for HW, Cc in ((4096, 320), (1024, 640), (1024, 320), (256, 1280), (256, 640), (64, 1280)):
    G = 32
    x = torch.randn(1, HW, Cc, device="cuda")
    gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    mean, rstd = torch.zeros(G, device="cuda"), torch.ones(G, device="cuda")
    tmp = torch.empty(nfl(HW, Cc, G, nb) + 64, device="cuda")
    ts = [torch.randn(nb, HW, Cc, device="cuda").half() for _ in range(4)]
    out = torch.empty(nb, HW, Cc, device="cuda", dtype=torch.float16)
    for mode in (0, 1):
        f = N.leaf("pbk_gn_lin")
        us = timeit(lambda i: f(p(x), p(mean), p(rstd), p(gamma), p(beta), HW, Cc, G, 1, p(ts[i % 4]), nb, mode, p(out), C.c_float(0), 6,
                                p(tmp), 0, C.c_long(0), st))
        byt = nb * HW * Cc * 4 + HW * Cc * 4                     # tangent in + out (halves), primal once
        res.append(("gn_lin", HW, Cc, mode, round(us, 1), round(byt / us / 1e3)))
for rows, Cc in ((4096, 320), (1024, 640), (256, 1280), (64, 1280)):
    x = torch.randn(rows, Cc, device="cuda")
    gamma = torch.randn(Cc, device="cuda")
    mean, rstd = torch.zeros(rows, device="cuda"), torch.ones(rows, device="cuda")
    ts = [torch.randn(nb, rows, Cc, device="cuda").half() for _ in range(4)]
    out = torch.empty(nb, rows, Cc, device="cuda", dtype=torch.float16)
    for mode in (0, 1):
        f = N.leaf("pbk_ln_lin")
        us = timeit(lambda i: f(p(x), p(mean), p(rstd), p(gamma), C.c_long(rows), Cc, p(ts[i % 4]), nb, mode, p(out), C.c_float(0), 6, 0,
                                C.c_long(0), st))
        res.append(("ln_lin", rows, Cc, mode, round(us, 1), round((nb * rows * Cc * 4 + rows * Cc * 4) / us / 1e3)))
    Fd = 4 * Cc
    h = torch.randn(rows, 2 * Fd, device="cuda")
    dh = [torch.randn(nb, rows, 2 * Fd, device="cuda").half() for _ in range(2)]
    dy = torch.empty(nb, rows, Fd, device="cuda", dtype=torch.float16)
    f = N.leaf("pbk_geglu_jvp")
    us = timeit(lambda i: f(p(h), C.c_long(rows), p(dh[i % 2]), nb, Fd, p(dy), 6, 0, C.c_long(0), st))
    res.append(("geglu_jvp", rows, Cc, 0, round(us, 1), round((nb * rows * Fd * 6 + rows * Fd * 8) / us / 1e3)))
    f = N.leaf("pbk_geglu_vjp")
    us = timeit(lambda i: f(p(h), C.c_long(rows), p(dy), nb, Fd, p(dh[i % 2]), 6, 0, C.c_long(0), st))
    res.append(("geglu_vjp", rows, Cc, 1, round(us, 1), round((nb * rows * Fd * 6 + rows * Fd * 8) / us / 1e3)))
print("kernel rows C mode us GB/s(algorithmic)")
for r in res:
    print(*r)
json.dump(res, open(os.path.join("gpurun_out", f"bench_elementwise_nb{nb}.json"), "w"))
