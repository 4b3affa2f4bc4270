"""Sanity at the reference's DEFAULT rank (pca_rank = 50, chunk_size = 25, utils.py:722-725) on the full SD-v1.5 mid-block
problem: finite, orthonormal right vectors, descending singular values, and ||J v_i|| = s_i (Rayleigh check through pb_jvp)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusion_pullback_b200 as PB
from diffusion_pullback_b200 import synthetic as SY
dev = torch.device("cuda:0")
unet = PB.patch_unet(SY.SyntheticUNet("sd15", upto=("mid", 0), device=dev))
x, t, ctx = SY.synthetic_inputs("sd15", device=dev)
torch.manual_seed(0)
t0 = time.time()
u, s, vT, info = unet.local_encoder_pullback_zt(x, t, ctx, op="mid", block_idx=0, pca_rank=50, min_iter=4, max_iter=4,
                                                convergence_threshold=0.0, return_info=True)
torch.cuda.synchronize()
dt = time.time() - t0
eng = next(iter(unet._pb200_engines.values()))
U = eng.jvp(vT)
ok = bool(torch.isfinite(s).all() and torch.isfinite(vT).all())
orth = float((vT @ vT.T - torch.eye(50, device=dev)).abs().max())
desc = bool((s[:-1] >= s[1:] - 1e-6).all())
ray = float(((U.norm(dim=1) - s).abs() / s).max())
print(dict(k=50, iters=info["iters_done"], seconds=round(dt, 2), finite=ok, orth_err=orth, descending=desc, rayleigh_rel=ray,
           s_top=[round(float(v), 3) for v in s[:5]]))
assert ok and orth < 1e-4 and desc
