"""Profiling driver (run under ncu on the GPU box): one problem of a bench workload, eager launches (no graph),
`--iters` subspace iterations.  See profiles/README.md for the commands that produced the committed summaries.

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python scripts/profile_iter.py --workload sd15_mid_k5_i50 --iters 2
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusion_pullback_b200 as PB                      # noqa: E402
from diffusion_pullback_b200 import synthetic as SY       # noqa: E402
from bench import WORKLOADS                               # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="sd15_mid_k5_i50")
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--graph", type=int, default=0)
ap.add_argument("--slots", type=int, default=1, help="problem slots (the bench default is 5 at k = 5)")
a = ap.parse_args()
model, op, bi, k, _, _, _ = WORKLOADS[a.workload]
dev = torch.device("cuda:0")
unet = SY.SyntheticUNet(model, upto=(op, bi), device=dev)
size = unet.config["sample_size"]
P = a.slots
eng = PB.PullbackEngine(PB.unet_config(unet), size, size, op, bi, k * P, unet.config["ctx_len"], dev)
eng.bind(unet.state_dict())
if P > 1:
    eng.set_slots(P)
eng.set_option("use_graph", a.graph)
x, t, ctx = SY.synthetic_inputs(model)
torch.manual_seed(0)
q = torch.cat([torch.linalg.qr(torch.randn(eng.n_in, k))[0].T for _ in range(P)]).contiguous()
torch.cuda.synchronize()
l0 = eng.launches
for sl in range(P):
    eng.set_point(x + 0.01 * sl, float(t), ctx, slot=sl)
l1 = eng.launches
u, s, vT, info = eng.pullback(q, a.iters, a.iters, 0.0)
torch.cuda.synchronize()
print("leaf calls: primal", l1 - l0, "iterations", eng.launches - l1, "s =", s.tolist())
