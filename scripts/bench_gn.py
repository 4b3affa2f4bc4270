"""Timing of the GroupNorm linearisation (pbk_gn_lin) on the SD-1.5 layer shapes, CUDA events; PB_GN_SPLIT selects the cluster size."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_pullback_b200 import _native as N
f = N.leaf("pbk_gn_lin")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
nfl = N.raw().pbk_gn_tmp_floats
nfl.restype = C.c_size_t
res = []
for nb, HW, Cc in ((5, 4096, 320), (5, 1024, 640), (5, 256, 1280), (5, 64, 1280)):
    G = 32
    x = torch.randn(1, HW, Cc, device="cuda")
    gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    mean, rstd = torch.zeros(G, device="cuda"), torch.ones(G, device="cuda")
    tmp = torch.empty(nfl(HW, Cc, G, nb) + 64, device="cuda")
    ts = [torch.randn(nb, HW, Cc, device="cuda") for _ in range(4)]
    out = torch.empty(nb, HW, Cc, device="cuda")
    for mode in (0, 1):
        for _ in range(3):
            f(p(x), p(mean), p(rstd), p(gamma), p(beta), HW, Cc, G, 1, p(ts[0]), nb, mode, p(out), C.c_float(0), 1, p(tmp), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            f(p(x), p(mean), p(rstd), p(gamma), p(beta), HW, Cc, G, 1, p(ts[i % 4]), nb, mode, p(out), C.c_float(0), 1, p(tmp), st)
        e1.record(); torch.cuda.synchronize()
        res.append((HW, Cc, mode, round(e0.elapsed_time(e1) * 50, 1)))
print("PB_GN_SPLIT", os.environ.get("PB_GN_SPLIT"), res)
