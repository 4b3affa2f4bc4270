"""Problem slots (pb_set_slots) on the GPU box: P independent problems batched through one handle against the same problems
solved one at a time -- per-problem agreement (singular values, subspace overlap) and throughput (iters/s, CUDA events).
    python scripts/bench_slots.py [--workload sd15_mid_k5_i50] [--slots 5] [--iters 50]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusion_pullback_b200 as PB
from diffusion_pullback_b200 import synthetic as SY

WORKLOADS = {"sd15_mid_k5_i50": ("sd15", "mid", 0, 5), "sd_small_mid_k4": ("sd_small", "mid", 0, 4), "sd15_up1_k5_i50": ("sd15", "up", 1, 5)}
ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="sd15_mid_k5_i50")
ap.add_argument("--slots", type=int, default=5)
ap.add_argument("--iters", type=int, default=50)
args = ap.parse_args()
name, op, bi, k = WORKLOADS[args.workload]
P, iters = args.slots, args.iters
dev = torch.device("cuda:0")
unet = SY.SyntheticUNet(name, upto=(op, bi), device=dev)
cfg = PB.unet_config(unet)
size, ctx_len = unet.config["sample_size"], unet.config["ctx_len"]
_, t, ctx = SY.synthetic_inputs(name)
sd = unet.state_dict()
eng1 = PB.PullbackEngine(cfg, size, size, op, bi, k, ctx_len, dev)
eng1.bind(sd)
engP = PB.PullbackEngine(cfg, size, size, op, bi, P * k, ctx_len, dev)
engP.bind(sd)
engP.set_slots(P)
g = torch.Generator().manual_seed(7)
xs = [torch.randn(1, cfg["in_channels"], size, size, generator=g).to(dev) for _ in range(P)]
ts = [float(t) - 37.0 * p for p in range(P)]
cs = [(ctx if p == 0 else torch.randn(ctx.shape, generator=g)).to(dev) for p in range(P)] if ctx is not None else [None] * P
v0 = []
for p in range(P):
    q, _ = torch.linalg.qr(torch.randn(eng1.n_in, k, generator=g))
    v0.append(q.T.contiguous().to(dev))
V0 = torch.cat(v0, 0)


def single():
    out = []
    for p in range(P):
        eng1.set_point(xs[p], ts[p], cs[p])
        out.append(eng1.pullback(v0[p], iters, iters, 0.0))
    return out


def batched():
    for p in range(P):
        engP.set_point(xs[p], ts[p], cs[p], slot=p)
    return engP.pullback(V0, iters, iters, 0.0)


def timed(fn):
    fn()                                                    # warm-up (graph capture)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / 1e3


res1, t1 = timed(single)
(u, s, vT, info), tP = timed(batched)
rep = []
for p in range(P):
    up, sp, vp, _ = res1[p]
    sl = slice(p * k, (p + 1) * k)
    s_rel = float(((s[sl] - sp).abs() / sp).max())
    overlap = float((vp.double() @ vT[sl].double().T).pow(2).sum() / k)
    rep.append({"s_rel_max": s_rel, "subspace": overlap})
print(json.dumps({"workload": args.workload, "slots": P, "k": k, "iters": iters,
                  "one_at_a_time_iters_per_s": P * iters / t1, "batched_iters_per_s": P * iters / tP, "speedup": t1 / tP,
                  "per_problem": rep}))
