"""Operand builders and fp64 references for pbk_attn_lin in the roles and layouts the ENGINE uses (run_attn_jvp / run_attn_vjp in
pb_engine.cpp): shared by tests/test_kernels_gpu.py and scripts/bench_attn.py.  Cases: "jvp" (two segments + folded P.C2 + row
sums), "jvp_qkv" (the same with the operands as column slices of [N][3C] / [nb][N][3C] tensors), "cross" (one segment + row sums),
"vjp_a" (row deltas), "vjp_b" (A primal, B per tangent, column deltas, separate D2)."""
import ctypes as C
import math

import torch

from diffusion_pullback_b200 import _native as N


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
    qkv_layout = case == "jvp_qkv"              # the engine's layout: operands are column slices of [N][3C] / [nb][N][3C] tensors
    if qkv_layout:
        case = "jvp"
        dqkv, qkv = hf(nb, Mr, 3 * Cc), hf(Mr, 3 * Cc)
        t["A0"], t["B1"] = dqkv[:, :, :Cc], dqkv[:, :, Cc:2 * Cc]
        t["A1"], t["B0"] = qkv[:, :Cc], qkv[:, Cc:2 * Cc]
        t["_keep"] = (dqkv, qkv)
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
    if qkv_layout:
        s0.A, s0.lda, s0.sAb, s0.sAh, s0.B, s0.ldb, s0.sBb, s0.sBh = dqkv.data_ptr(), 3 * Cc, Mr * 3 * Cc, d, qkv.data_ptr() + 2 * Cc, 3 * Cc, 0, d
        s1.A, s1.lda, s1.sAb, s1.sAh, s1.B, s1.ldb, s1.sBb, s1.sBh = qkv.data_ptr(), 3 * Cc, 0, d, dqkv.data_ptr() + 2 * Cc, 3 * Cc, Mr * 3 * Cc, d
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
    t["case"] = case
    nprod = nseg + 1 + (1 if c2 else 0)                          # contractions with the score matrix, 2 Mr Nc d flops each
    return a, t, nprod, scale


def reference(t, Mr, Nc, d, nb, nh, case, scale):
    Cc = nh * d
    case = "jvp" if case == "jvp_qkv" else case
    dd = lambda x: x.double().contiguous()
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


