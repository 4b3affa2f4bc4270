"""Timing of the SURVEY.md s.8f row-1 path on one B200: the whole SD-v1.5 U-Net forward (x_t -> eps, batch 1, FULL plan,
TF32 primal pass incl. the linearisation cache it fills) and the fused DDIM update, CUDA events.

    python scripts/bench_ddim.py [--model sd15] [--steps 10]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusion_pullback_b200 as PB                      # noqa: E402
from diffusion_pullback_b200 import synthetic as SY       # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="sd15")
ap.add_argument("--steps", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
unet = PB.patch_unet(SY.SyntheticUNet(a.model, device=dev))
z, t, ctx = SY.synthetic_inputs(a.model, device=dev)
betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000) ** 2
sched = PB.DDIMSchedule(torch.cumprod(1 - betas, 0))
for _ in range(3):
    e = unet.eps(z, t, ctx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    e = unet.eps(z, t, ctx)
e1.record()
torch.cuda.synchronize()
fwd_ms = e0.elapsed_time(e1) / a.steps
sched.set_timesteps(50)
ts = list(sched.timesteps)
e0.record()
for i in range(a.steps):
    z2 = sched.step(e, ts[i % len(ts)], z).prev_sample
e1.record()
torch.cuda.synchronize()
step_us = 1e3 * e0.elapsed_time(e1) / a.steps
e0.record()
zT = PB.ddim_forward_steps(unet, sched, z, ctx, 11)          # 10 U-Net calls + 10 updates
e1.record()
torch.cuda.synchronize()
print(json.dumps({"model": a.model, "unet_forward_ms": fwd_ms, "unet_forwards_per_s": 1e3 / fwd_ms, "ddim_step_us": step_us,
                  "ddim_10_steps_ms": e0.elapsed_time(e1), "primal_gflop": 261.4 if a.model == "sd15" else None,
                  "note": "full forward = the engine's primal pass over the FULL plan (TF32, fills the linearisation cache: P, P^T and "
                          "the transposed operands of every attention layer); algorithmic flops of the full SD-1.5 forward ~ 0.8 TF"}))
