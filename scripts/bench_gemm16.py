"""fp16-operand GEMM shapes of the batched iteration (5 problem slots x k = 5, SD-1.5 mid-block), one-CTA kernel against the
CTA-pair kernel (tcgen05.mma.cta_group::2): CUDA events over 20 launches, operands rotated through > 126 MB (GPU box).
  python scripts/bench_gemm16.py [nb [shape indices [0 | 1: one kernel only]]]
"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from diffusion_pullback_b200 import _native as N

nbat = int(sys.argv[1]) if len(sys.argv) > 1 else 25
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
f = N.leaf("pbk_gemm")
pair = N.raw().pb_gemm_tune_pair
ws = torch.empty(16 << 20, device="cuda")

# (conv, H, W, Cin / K, Cout / N, residual)
shapes = [(1, 64, 64, 320, 320, 0), (1, 32, 32, 640, 640, 0), (1, 16, 16, 1280, 1280, 0), (1, 8, 8, 1280, 1280, 0),
          (1, 32, 32, 320, 640, 0), (1, 16, 16, 640, 1280, 0), (1, 16, 16, 1280, 640, 0),
          (0, 64, 64, 320, 320, 0), (0, 64, 64, 320, 320, 1), (0, 64, 64, 320, 2560, 0), (0, 64, 64, 1280, 320, 1),
          (0, 64, 64, 320, 960, 0), (0, 32, 32, 640, 640, 0), (0, 32, 32, 640, 5120, 0), (0, 32, 32, 2560, 640, 1),
          (0, 16, 16, 1280, 1280, 0), (0, 16, 16, 1280, 10240, 0), (0, 16, 16, 5120, 1280, 1), (0, 8, 8, 1280, 1280, 0)]
if len(sys.argv) > 2:
    shapes = [shapes[int(i)] for i in sys.argv[2].split(",")]
print(f"nb = {nbat}; us per launch and TF/s: one-CTA | CTA pair")
for conv, H, W, K, Nn, res in shapes:
    M = nbat * H * W
    flops = 2.0 * M * Nn * (9 * K if conv else K)
    nrot = max(2, int(200e6 // (M * (K + Nn) * 2)) + 1)
    A = torch.randn(nrot, M, K, device="cuda").half()
    B = (torch.randn(nrot, Nn, 9 * K if conv else K, device="cuda") * 0.05).half()
    D = torch.empty(nrot, M, Nn, device="cuda", dtype=torch.float16)
    gs = []
    for r in range(nrot):
        g = N.PbGemm()
        g.M, g.N, g.nseg = M, Nn, 1
        s = g.seg[0]
        s.A, s.lda, s.B, s.ldb, s.K = A[r].data_ptr(), K, B[r].data_ptr(), B.shape[-1], K
        g.D, g.ldd = D[r].data_ptr(), Nn
        if res:
            g.R, g.ldr, g.beta = D[r].data_ptr(), Nn, 1.0
        g.alpha, g.nb, g.nh, g.conv, g.H, g.W = 1.0, (nbat if conv else 1), 1, conv, H, W
        g.ab_dtype, g.d_dtype = 1, 1
        g.ws, g.ws_floats = ws.data_ptr(), ws.numel()
        gs.append(g)
    out = []
    ref = None
    for on in ((0, 1) if len(sys.argv) < 4 else (int(sys.argv[3]),)):
        pair(on)
        for g in gs:
            err = f(C.byref(g), st)
            assert err is None, err
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            f(C.byref(gs[i % nrot]), st)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 20
        out.append(us)
        if not res:
            if ref is None:
                ref = D[0].clone()
            else:
                assert (D[0].float() - ref.float()).abs().max() <= 1e-2 * ref.float().abs().max(), "pair kernel differs"
    pair(1)
    if len(out) == 1:
        out = out * 2
    print(f"{'conv3x3' if conv else 'linear '} M={M:6d} N={Nn:5d} K={K:5d} res={res}:  {out[0]:8.1f} us {flops / out[0] / 1e6:7.1f} | "
          f"{out[1]:8.1f} us {flops / out[1] / 1e6:7.1f}   x{out[0] / out[1]:.2f}")
