"""One launch of each fp16-tangent elementwise kernel on its largest SD-1.5 shapes, for an `ncu --set full` capture (profiles/):
    ncu --set full --clock-control none -o gpurun_out/elem python scripts/ncu_elementwise.py --nb 25"""
import argparse, ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_pullback_b200 import _native as N
ap = argparse.ArgumentParser(); ap.add_argument("--nb", type=int, default=25); a = ap.parse_args()
nb = a.nb
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
nfl = N.raw().pbk_gn_tmp_floats
nfl.restype = C.c_size_t
for HW, Cc in ((4096, 320), (1024, 640)):
    G = 32
    x = torch.randn(1, HW, Cc, device="cuda"); gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    mean, rstd = torch.zeros(G, device="cuda"), torch.ones(G, device="cuda")
    tmp = torch.empty(nfl(HW, Cc, G, nb) + 64, device="cuda")
    t = torch.randn(nb, HW, Cc, device="cuda").half(); out = torch.empty_like(t)
    for mode in (0, 1):
        N.leaf("pbk_gn_lin")(p(x), p(mean), p(rstd), p(gamma), p(beta), HW, Cc, G, 1, p(t), nb, mode, p(out), C.c_float(0), 6, p(tmp), 0, C.c_long(0), st)
rows, Cc = 4096, 320
x = torch.randn(rows, Cc, device="cuda"); gamma = torch.randn(Cc, device="cuda")
mean, rstd = torch.zeros(rows, device="cuda"), torch.ones(rows, device="cuda")
t = torch.randn(nb, rows, Cc, device="cuda").half(); out = torch.empty_like(t)
for mode in (0, 1):
    N.leaf("pbk_ln_lin")(p(x), p(mean), p(rstd), p(gamma), C.c_long(rows), Cc, p(t), nb, mode, p(out), C.c_float(0), 6, 0, C.c_long(0), st)
Fd = 4 * Cc
h = torch.randn(rows, 2 * Fd, device="cuda"); dh = torch.randn(nb, rows, 2 * Fd, device="cuda").half()
dy = torch.empty(nb, rows, Fd, device="cuda", dtype=torch.float16)
N.leaf("pbk_geglu_jvp")(p(h), C.c_long(rows), p(dh), nb, Fd, p(dy), 6, 0, C.c_long(0), st)
N.leaf("pbk_geglu_vjp")(p(h), C.c_long(rows), p(dy), nb, Fd, p(dh), 6, 0, C.c_long(0), st)
# attn_delta, transposes at the 4096-token layer
Hh, d = 8, 40
go = torch.randn(nb, rows, Cc, device="cuda").half(); o = torch.randn(rows, Cc, device="cuda")
delta = torch.empty(nb, Hh, rows, device="cuda")
N.leaf("pbk_attn_delta")(p(go), C.c_long(Cc), p(o), C.c_long(Cc), nb, rows, Hh, d, p(delta), 4, 0, C.c_long(0), st)
dst = torch.empty(nb, Hh, d, rows, device="cuda", dtype=torch.float16)
N.leaf("pbk_transpose")(p(dst), C.c_long(rows), C.c_long(Cc * rows), C.c_long(d * rows), p(go), C.c_long(Cc), C.c_long(rows * Cc), C.c_long(d), nb, Hh, rows, d,
                        C.c_float(0), 6, st)
# re-orthonormalisation at k = 25 columns of one 16384-element latent (5 slots run it per problem at k = 5)
k, n = 5, 16384
W = torch.randn(k, n, device="cuda"); Vp = torch.linalg.qr(torch.randn(n, k, device="cuda"))[0].T.contiguous()
G_ = torch.zeros(k * k, device="cuda", dtype=torch.float64); M_ = torch.zeros(k * k, device="cuda", dtype=torch.float64)
Rm = torch.empty(k * k, device="cuda"); sv = torch.empty(k, device="cuda"); V = torch.empty(k, n, device="cuda"); met = torch.zeros(4, device="cuda")
N.leaf("pbk_gram2")(p(W), p(Vp), k, C.c_long(n), p(G_), p(M_), st)
N.leaf("pbk_jacobi")(p(G_), p(M_), k, p(Rm), p(sv), st)
N.leaf("pbk_rotate")(p(W), p(Rm), p(Vp), k, C.c_long(n), C.c_float(0), C.c_float(0), p(V), p(met), st)
torch.cuda.synchronize()
print("done")
