"""Timing of the fused attention-linearisation kernel (pbk_attn_lin, all-fp16 plan) alone, in the four roles the engine uses,
on the SD layer shapes: CUDA events over back-to-back launches.  PB_ATTN_V1=1 selects the per-column kernel
(pb_attn_sm100.cu) instead of the column-batched one (pb_attn16_sm100.cu); PB_ATTN_KC caps the columns per CTA.

    python scripts/bench_attn.py [--nb 5] [--shapes sd15|sd21|all] [--check]
"""
import argparse, ctypes as C, json, math, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusion_pullback_b200 import _native as N

ap = argparse.ArgumentParser()
ap.add_argument("--nb", type=int, default=5)
ap.add_argument("--shapes", default="sd15")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--check", action="store_true", help="compare with an fp64 torch evaluation (small shapes only)")
a_ = ap.parse_args()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)


from tests.attn_cases import build, reference


rel = lambda x, y: float((x.double() - y.double()).norm() / y.double().norm())
shapes = []
if a_.shapes in ("sd15", "all"):
    shapes += [(4096, 4096, 40, 8, c) for c in ("jvp", "vjp_a", "vjp_b")] + [(4096, 77, 40, 8, "cross"), (4096, 77, 40, 8, "vjp_a")]
if a_.shapes in ("sd21", "all"):
    shapes += [(9216, 9216, 64, 5, c) for c in ("jvp", "vjp_a", "vjp_b")]
if a_.shapes.startswith("one:"):
    c_ = a_.shapes[4:]
    n_ = int(os.environ.get("BENCH_N", "4096"))
    shapes += [(n_, 77 if c_ == "cross" else n_, int(os.environ.get("BENCH_D", "40")), int(os.environ.get("BENCH_NH", "8")), c_)]
if a_.shapes == "qkv":
    shapes += [(1024, 1024, 16, 4, "jvp_qkv"), (1024, 1024, 40, 2, "jvp_qkv"), (512, 512, 64, 2, "jvp_qkv")]
if a_.shapes == "d16":
    shapes += [(1024, 1024, 16, 4, c) for c in ("jvp", "vjp_a", "vjp_b", "cross")] + [(1024, 1024, 32, 4, c) for c in ("jvp", "vjp_a", "vjp_b")]
if a_.shapes == "small":
    shapes += [(512, 512, 40, 2, c) for c in ("jvp", "vjp_a", "vjp_b", "cross")] + [(384, 320, 64, 2, c) for c in ("jvp", "vjp_a", "vjp_b")]
res = []
f = N.leaf("pbk_attn_lin")
for Mr, Nc, d, nh, case in shapes:
    a, t, nprod, scale = build(Mr, Nc, d, a_.nb, nh, case)
    for _ in range(2):
        err = f(C.byref(a), st)
        assert not err, err
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a_.reps):
        f(C.byref(a), st)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / a_.reps
    tf = nprod * 2.0 * Mr * Nc * d * nh * a_.nb / us / 1e6
    row = dict(case=case, Mr=Mr, Nc=Nc, d=d, nh=nh, nb=a_.nb, us=round(us, 1), tflops=round(tf, 1))
    if a_.check:
        ref, e2 = reference(t, Mr, Nc, d, a_.nb, nh, case, scale)
        row["err"] = rel(t["D"].float(), ref)
        if e2 is not None:
            row["err2"] = rel(t["D2"].float(), e2)
    res.append(row)
    print(row, flush=True)
    del a, t
    torch.cuda.empty_cache()
tag = "v1" if os.environ.get("PB_ATTN_V1") else "v2"
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open(os.path.join("gpurun_out", f"bench_attn_{tag}_nb{a_.nb}_{a_.shapes.replace(':', '_')}.json"), "w"))
