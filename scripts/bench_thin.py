"""conv_in tangent kernels (pbk_conv3x3_direct: 4 -> 320 channels fp32 in / fp16 out, and the transpose fp16 in / fp32 out) at the
batched iteration's size (25 images of 64 x 64): CUDA events over 20 launches."""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from diffusion_pullback_b200 import _native as N
nb, H, W, Ci, Co = int(sys.argv[1]) if len(sys.argv) > 1 else 25, 64, 64, 4, 320
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr())
w = torch.randn(Co, Ci, 3, 3, device="cuda") / 6
fwd, bwd = torch.empty(Co, 9, Ci, device="cuda"), torch.empty(Ci, 9, Co, device="cuda")
N.leaf("pbk_pack_conv3x3")(p(w), Co, Ci, p(fwd), p(bwd), 0, st)
xi = torch.randn(nb, H, W, Ci, device="cuda")
yo = torch.empty(nb, H, W, Co, device="cuda", dtype=torch.float16)
go = torch.randn(nb, H, W, Co, device="cuda").half()
gi = torch.empty(nb, H, W, Ci, device="cuda")
f = N.leaf("pbk_conv3x3_direct")
for name, args in (("thin-in  (JVP)", (p(xi), nb, H, W, Ci, p(fwd), None, Co, p(yo), C.c_float(0), 2, st)),
                   ("thin-out (VJP)", (p(go), nb, H, W, Co, p(bwd), None, Ci, p(gi), C.c_float(0), 4, st))):
    for _ in range(3):
        assert f(*args) is None
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f(*args)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(f"{name}: {us:7.1f} us  ({2.0 * nb * H * W * Co * Ci * 9 / us / 1e6:.1f} TFLOP/s fp32, {nb * H * W * Co * 2 / us / 1e3:.0f} GB/s of the 320-channel tensor)")
