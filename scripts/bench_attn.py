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


def build(Mr, Nc, d, nb, nh, case):
    Cc = nh * d
    ldp = (Nc + 7) // 8 * 8
    scale = 2.0 ** round(0.5 * math.log2(Nc))
    hf = lambda *s: (torch.randn(*s, device="cuda") * 0.5).half()
    t = dict(A0=hf(nb, Mr, Cc), B0=hf(Nc, Cc), A1=hf(Mr, Cc), B1=hf(nb, Nc, Cc), C1=hf(nh, d, ldp), C2=hf(nb, nh, d, ldp),
             O=torch.randn(Mr, Cc, device="cuda"))
    P16 = torch.empty(nh, Mr, ldp, device="cuda", dtype=torch.float16)
    for h in range(nh):                                          # head by head: the fp32 softmax of 4096^2 x 8 is 0.5 GB
        Ph = torch.softmax(torch.randn(Mr, ldp, device="cuda") * 2, -1)
        Ph[:, Nc:] = 0
        P16[h] = (Ph * scale).half()
    t["P16"] = P16
    nseg = 2 if case == "jvp" else 1
    mode = {"jvp": 0, "cross": 0, "vjp_a": 1, "vjp_b": 2}[case]
    c2 = {"jvp": 1, "cross": 0, "vjp_a": 0, "vjp_b": 2}[case]
    t["delta"] = torch.randn(nb, nh, Mr if mode == 1 else Nc, device="cuda") if mode else None
    t["D"] = torch.zeros(nb, Mr, Cc, device="cuda", dtype=torch.float16)
    t["D2"] = torch.zeros(nb, Mr, Cc, device="cuda", dtype=torch.float16)
    a = N.PbAttnLin()
    a.Mr, a.Nc, a.d, a.nb, a.nh, a.nseg = Mr, Nc, d, nb, nh, nseg
    s0 = a.seg[0]
    if case == "vjp_b":      # A primal (V rows), B per tangent (Obar)
        s0.A, s0.lda, s0.sAb, s0.sAh, s0.B, s0.ldb, s0.sBb, s0.sBh = t["A1"].data_ptr(), Cc, 0, d, t["B1"].data_ptr(), Cc, Nc * Cc, d
    else:
        s0.A, s0.lda, s0.sAb, s0.sAh, s0.B, s0.ldb, s0.sBb, s0.sBh = t["A0"].data_ptr(), Cc, Mr * Cc, d, t["B0"].data_ptr(), Cc, 0, d
    s1 = a.seg[1]
    s1.A, s1.lda, s1.sAb, s1.sAh, s1.B, s1.ldb, s1.sBb, s1.sBh = t["A1"].data_ptr(), Cc, 0, d, t["B1"].data_ptr(), Cc, Nc * Cc, d
    a.alpha1, a.alpha2, a.beta = d ** -0.5, 0.7, 0.0
    a.Pm, a.ldp, a.sPh = P16.data_ptr(), ldp, Mr * ldp
    a.delta, a.delta_mode = (t["delta"].data_ptr() if mode else None), mode
    a.want_rsum, a.O, a.ldo = int(mode == 0), t["O"].data_ptr(), Cc
    a.C1, a.ldc, a.sCh = t["C1"].data_ptr(), ldp, d * ldp
    a.D, a.ldd, a.sDb, a.round_tf32 = t["D"].data_ptr(), Cc, Mr * Cc, 1
    a.p16, a.p_scale, a.s16 = 1, scale, 1
    if c2:
        a.C2, a.ldc2, a.sC2h, a.sC2b = t["C2"].data_ptr(), ldp, d * ldp, nh * d * ldp
    if c2 == 2:
        a.D2, a.ldd2, a.sD2b = t["D2"].data_ptr(), Cc, Mr * Cc
    nprod = nseg + 1 + (1 if c2 else 0)                          # contractions with the score matrix, 2 Mr Nc d flops each
    return a, t, nprod, scale


def reference(t, Mr, Nc, d, nb, nh, case, scale):
    Cc = nh * d
    dd = lambda x: x.double()
    if case == "vjp_b":
        S = torch.einsum("ihd,bjhd->bhij", dd(t["A1"]).view(Mr, nh, d), dd(t["B1"]).view(nb, Nc, nh, d))
    else:
        S = torch.einsum("bihd,jhd->bhij", dd(t["A0"]).view(nb, Mr, nh, d), dd(t["B0"]).view(Nc, nh, d))
    if case == "jvp":
        S = S + torch.einsum("ihd,bjhd->bhij", dd(t["A1"]).view(Mr, nh, d), dd(t["B1"]).view(nb, Nc, nh, d))
    S = S * d ** -0.5
    if case == "vjp_a":
        S = S - dd(t["delta"])[..., :, None]
    if case == "vjp_b":
        S = S - dd(t["delta"])[..., None, :]
    Ps = dd(t["P16"])[None, :, :, :Nc]
    Tr = (Ps * S).float().half().double()
    acc = torch.einsum("bhij,hnj->bihn", Tr, dd(t["C1"])[..., :Nc]).reshape(nb, Mr, Cc) / scale
    e2 = None
    if case in ("jvp", "vjp_b"):
        e2 = torch.einsum("hij,bhnj->bihn", dd(t["P16"])[..., :Nc], dd(t["C2"])[..., :Nc]).reshape(nb, Mr, Cc) / scale
        if case == "jvp":
            acc = acc + e2
    ref = 0.7 * acc
    if case in ("jvp", "cross"):
        rs = Tr.sum(-1) / scale
        ref = ref - (rs.permute(0, 2, 1)[..., None] * dd(t["O"]).view(Mr, nh, d)[None]).reshape(nb, Mr, Cc)
    return ref, (e2 if case == "vjp_b" else None)


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
