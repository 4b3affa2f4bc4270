"""GPU unit tests of the leaf kernels (pbk_*), each against plain PyTorch fp32/fp64 of the same op.
The tcgen05 GEMM multiplies TF32-truncated operands with fp32 accumulation; the reference for it is
an fp64 product of the same truncated operands, so the tolerance only covers accumulation order."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from diffusion_pullback_b200 import _native as N

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False          # the torch references must be true fp32
torch.backends.cuda.matmul.allow_tf32 = False


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _ok(err):
    assert err is None, err.decode()
    torch.cuda.synchronize()


def tf32_trunc(t):
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def gemm(A, B, D, *, M, N_, K, lda, ldb, ldd, sAb=0, sAh=0, sBb=0, sBh=0, sDb=0, sDh=0, nb=1, nh=1,
         bias=None, R=None, ldr=0, sRb=0, sRh=0, alpha=1.0, beta=0.0, conv=0, H=0, W=0, seg2=None, rnd=0, splitk=True,
         ab_dtype=0, d_dtype=0):
    g = N.PbGemm()
    g.M, g.N, g.nseg = M, N_, 1 if seg2 is None else 2
    s = g.seg[0]
    s.A, s.lda, s.sAb, s.sAh, s.B, s.ldb, s.sBb, s.sBh, s.K = A.data_ptr(), lda, sAb, sAh, B.data_ptr(), ldb, sBb, sBh, K
    if seg2 is not None:
        A2, B2, K2, lda2, ldb2 = seg2
        s = g.seg[1]
        s.A, s.lda, s.sAb, s.sAh, s.B, s.ldb, s.sBb, s.sBh, s.K = A2.data_ptr(), lda2, 0, 0, B2.data_ptr(), ldb2, 0, 0, K2
    g.D, g.ldd, g.sDb, g.sDh = D.data_ptr(), ldd, sDb, sDh
    g.R = R.data_ptr() if R is not None else None
    g.ldr, g.sRb, g.sRh = ldr, sRb, sRh
    g.bias = bias.data_ptr() if bias is not None else None
    g.alpha, g.beta, g.nb, g.nh, g.conv, g.H, g.W, g.round_tf32 = alpha, beta, nb, nh, conv, H, W, rnd
    g.ab_dtype, g.d_dtype = ab_dtype, d_dtype
    if splitk:                               # split-K scratch as the engine provides it
        ws = _scratch()
        g.ws, g.ws_floats = ws.data_ptr(), ws.numel()
    _ok(N.leaf("pbk_gemm")(C.byref(g), _st()))


_SCRATCH = []


def _scratch():
    if not _SCRATCH:
        _SCRATCH.append(torch.empty(8 << 20, device="cuda"))
    return _SCRATCH[0]


def gemm16(A, B, D, *, M, N_, K, lda, ldb, ldd, out16=True, **kw):
    """fp16-operand GEMM (kind::f16); D / R are fp16 (out16) or fp32."""
    return gemm(A, B, D, M=M, N_=N_, K=K, lda=lda, ldb=ldb, ldd=ldd, ab_dtype=1, d_dtype=1 if out16 else 0, **kw)


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("M,Nn,K", [(128, 128, 32), (256, 64, 64), (300, 200, 96), (64, 40, 40), (1000, 1280, 1280),
                                    (320, 2560, 1280), (77, 320, 768), (4096, 320, 320), (5, 64, 1000), (320, 1280, 11520),
                                    (20480, 320, 320), (130, 7, 64), (257, 13, 100), (64, 16, 16), (1280, 1920, 640),
                                    (700, 77, 160), (3000, 4096, 40)])
def test_gemm_plain(M, Nn, K):
    torch.manual_seed(M + Nn + K)
    Kp = (K + 3) // 4 * 4
    A = torch.randn(M, Kp, device="cuda")
    B = torch.randn(Nn, Kp, device="cuda")
    Np = (Nn + 3) // 4 * 4
    bias = torch.randn(Np, device="cuda")
    R = torch.randn(M, Np, device="cuda")
    D = torch.full((M, Np), float("nan"), device="cuda")
    gemm(A, B, D, M=M, N_=Nn, K=K, lda=Kp, ldb=Kp, ldd=Np, bias=bias, R=R, ldr=Np, alpha=0.5, beta=2.0)
    ref = 0.5 * (tf32_trunc(A)[:, :K].double() @ tf32_trunc(B)[:, :K].double().T) + bias[:Nn].double() + 2.0 * R[:, :Nn].double()
    assert rel(D[:, :Nn], ref) < 1e-5
    if Np > Nn:
        assert torch.isnan(D[:, Nn:]).all()           # columns beyond N untouched


@pytest.mark.parametrize("out16", [True, False])
@pytest.mark.parametrize("M,Nn,K", [(128, 128, 64), (300, 200, 96), (64, 40, 40), (1000, 1280, 1280), (320, 1280, 11520),
                                    (20480, 320, 320), (130, 7, 64), (257, 13, 104), (64, 16, 16), (700, 77, 160),
                                    (3000, 4096, 40), (4096, 960, 320)])
def test_gemm_f16_plain(M, Nn, K, out16):
    torch.manual_seed(M + Nn + K)
    Kp = (K + 7) // 8 * 8
    A = torch.randn(M, Kp, device="cuda").half()
    B = torch.randn(Nn, Kp, device="cuda").half()
    q = 8 if out16 else 4
    Np = (Nn + q - 1) // q * q
    odt = torch.float16 if out16 else torch.float32
    bias = torch.randn(Np, device="cuda")
    R = torch.randn(M, Np, device="cuda").to(odt)
    D = torch.full((M, Np), float("nan"), device="cuda", dtype=odt)
    gemm16(A, B, D, M=M, N_=Nn, K=K, lda=Kp, ldb=Kp, ldd=Np, out16=out16, bias=bias, R=R, ldr=Np, alpha=0.25, beta=2.0)
    ref = 0.25 * (A[:, :K].double() @ B[:, :K].double().T) + bias[:Nn].double() + 2.0 * R[:, :Nn].double()
    assert rel(D[:, :Nn], ref) < (6e-4 if out16 else 1e-5)
    if out16:                                           # correctly rounded: at most one fp16 ulp from the exact result
        assert ((D[:, :Nn].double() - ref).abs() <= ref.abs() * 2 ** -10 + 1e-3).all()
    if Np > Nn:
        assert torch.isnan(D[:, Nn:]).all()


@pytest.mark.parametrize("nb,H,W,Ci,Co", [(1, 64, 64, 320, 320), (5, 8, 8, 1280, 1280), (3, 16, 16, 64, 96), (2, 32, 32, 640, 320),
                                          (5, 4, 4, 32, 64), (1, 2, 2, 32, 32), (2, 12, 12, 96, 64)])
def test_gemm_f16_conv3x3(nb, H, W, Ci, Co):
    """fp16 implicit-GEMM conv: filter read as a (channel, tap, out) tensor, so channel counts need not fill a k-block."""
    torch.manual_seed(2)
    x = torch.randn(nb, H, W, Ci, device="cuda").half()
    w = (torch.randn(Co, Ci, 3, 3, device="cuda") / math.sqrt(9 * Ci)).half()
    fwd = w.permute(0, 2, 3, 1).reshape(Co, 9 * Ci).contiguous()
    R = torch.randn(nb, H, W, Co, device="cuda").half()
    y = torch.zeros(nb, H, W, Co, device="cuda", dtype=torch.float16)
    gemm16(x, fwd, y, M=nb * H * W, N_=Co, K=Ci, lda=Ci, ldb=9 * Ci, ldd=Co, nb=nb, conv=1, H=H, W=W, R=R, ldr=Co, beta=1.0)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), None, padding=1).permute(0, 2, 3, 1) + R.double()
    assert rel(y, ref) < 6e-4


def _pair(on):
    N.raw().pb_gemm_tune_pair(int(on))


@pytest.mark.parametrize("out16", [True, False])
@pytest.mark.parametrize("M,Nn,K,res", [(20480, 320, 320, True), (19000, 320, 320, False), (19001, 640, 640, True),
                                        (12800, 1280, 320, False), (9500, 960, 328, True), (9600, 2560, 72, False),
                                        (128 * 75, 1280, 1280, True), (128 * 149, 160, 2560, False)])
def test_gemm_f16_cta_pair_plain(M, Nn, K, res, out16):
    """CTA-pair kernel (tcgen05.mma.cta_group::2, 256-row tiles over two SMs) against fp64 and, without split-K, bitwise
    against the one-CTA kernel (same k order); odd row-tile counts exercise the phantom half of the last pair."""
    torch.manual_seed(M + Nn + K)
    Kp = (K + 7) // 8 * 8
    A = torch.randn(M, Kp, device="cuda").half()
    B = (torch.randn(Nn, Kp, device="cuda") / math.sqrt(K)).half()
    odt = torch.float16 if out16 else torch.float32
    bias = torch.randn(Nn, device="cuda")
    R = torch.randn(M, Nn, device="cuda").to(odt) if res else None
    outs = []
    for on, splitk in ((1, True), (1, False), (0, False)):
        _pair(on)
        D = torch.full((M, Nn), float("nan"), device="cuda", dtype=odt)
        gemm16(A, B, D, M=M, N_=Nn, K=K, lda=Kp, ldb=Kp, ldd=Nn, out16=out16, bias=bias, R=R, ldr=Nn, alpha=0.5,
               beta=2.0 if res else 0.0, splitk=splitk)
        outs.append(D)
    _pair(1)
    ref = 0.5 * (A[:, :K].double() @ B[:, :K].double().T) + bias.double() + (2.0 * R.double() if res else 0.0)
    for D in outs:
        assert rel(D, ref) < (6e-4 if out16 else 1e-5)
    assert torch.equal(outs[1], outs[2])


@pytest.mark.parametrize("nb,H,W,Ci,Co", [(5, 64, 64, 320, 320), (25, 16, 16, 640, 1280), (7, 32, 32, 64, 640), (3, 48, 40, 96, 160),
                                          (25, 32, 32, 320, 640)])
def test_gemm_f16_cta_pair_conv3x3(nb, H, W, Ci, Co):
    torch.manual_seed(3)
    x = torch.randn(nb, H, W, Ci, device="cuda").half()
    w = (torch.randn(Co, Ci, 3, 3, device="cuda") / math.sqrt(9 * Ci)).half()
    fwd = w.permute(0, 2, 3, 1).reshape(Co, 9 * Ci).contiguous()
    R = torch.randn(nb, H, W, Co, device="cuda").half()
    outs = []
    for on, splitk in ((1, True), (1, False), (0, False)):
        _pair(on)
        y = torch.zeros(nb, H, W, Co, device="cuda", dtype=torch.float16)
        gemm16(x, fwd, y, M=nb * H * W, N_=Co, K=Ci, lda=Ci, ldb=9 * Ci, ldd=Co, nb=nb, conv=1, H=H, W=W, R=R, ldr=Co, beta=1.0,
               splitk=splitk)
        outs.append(y)
    _pair(1)
    ref = F.conv2d(x.double().permute(0, 3, 1, 2), w.double(), None, padding=1).permute(0, 2, 3, 1) + R.double()
    for y in outs:
        assert rel(y, ref) < 6e-4
    assert torch.equal(outs[1], outs[2])


@pytest.mark.parametrize("P,k,rows_p,Cc", [(1, 3, 1024, 320), (2, 2, 700, 640), (3, 1, 256, 160 * 8)])
def test_gemm_geglu_tangent_epilogue(P, k, rows_p, Cc):
    """ff1 tangent GEMM with the GEGLU linearisation in its epilogue (PbGemm::gg: interleaved weight rows, factor cache per problem)
    against the two separate kernels (GEMM -> pbk_geglu_jvp) and fp64."""
    torch.manual_seed(P + k + Cc)
    Fd = 4 * Cc if Cc <= 640 else 1280
    nb = P * k
    M = nb * rows_p
    w = (torch.randn(2 * Fd, Cc, device="cuda") / math.sqrt(Cc)).half()
    wi = torch.empty_like(w)
    _ok(N.leaf("pbk_interleave_rows16")(_p(wi), _p(w), Fd, Cc, _st()))
    ref_i = torch.stack([w[:Fd].view(Fd // 32, 32, Cc), w[Fd:].view(Fd // 32, 32, Cc)], 1).reshape(2 * Fd, Cc)
    assert torch.equal(wi, ref_i)
    ps = rows_p * 2 * Fd + 64                                        # floats between the problems' factor caches
    hp = torch.randn(P, ps, device="cuda")
    dx = torch.randn(M, Cc, device="cuda").half()
    # separate path
    dh = torch.empty(M, 2 * Fd, device="cuda", dtype=torch.float16)
    gemm16(dx, w, dh, M=M, N_=2 * Fd, K=Cc, lda=Cc, ldb=Cc, ldd=2 * Fd)
    sep = torch.empty(M, Fd, device="cuda", dtype=torch.float16)
    _ok(N.leaf("pbk_geglu_jvp")(_p(hp), C.c_long(rows_p), _p(dh), nb, Fd, _p(sep), 2 | 4, k if P > 1 else 0, C.c_long(ps if P > 1 else 0), _st()))
    # fused
    g = N.PbGemm()
    g.M, g.N, g.nseg = M, 2 * Fd, 1
    sg = g.seg[0]
    sg.A, sg.lda, sg.B, sg.ldb, sg.K = dx.data_ptr(), Cc, wi.data_ptr(), Cc, Cc
    out = torch.full((M, Fd), float("nan"), device="cuda", dtype=torch.float16)
    g.D, g.ldd, g.alpha, g.nb, g.nh, g.ab_dtype, g.d_dtype = out.data_ptr(), Fd, 1.0, 1, 1, 1, 1
    g.gg, g.gg_F, g.gg_rows_p, g.gg_k_slot, g.gg_p_stride = hp.data_ptr(), Fd, rows_p, (k if P > 1 else 0), (ps if P > 1 else 0)
    _ok(N.leaf("pbk_gemm")(C.byref(g), _st()))
    fac = hp[:, :rows_p * 2 * Fd].view(P, 1, rows_p, 2 * Fd).expand(P, k, rows_p, 2 * Fd).reshape(M, 2 * Fd).double()
    acc = dx.double() @ w.double().T
    ref = acc[:, :Fd] * fac[:, :Fd] + acc[:, Fd:] * fac[:, Fd:]
    assert rel(out, ref) < 6e-4
    assert rel(out.float(), sep.float()) < 1e-3                     # the separate path rounds dh to halves in between


def test_gemm_f16_attention_batched():
    """Head-strided fp16 operands, fp32 scores out, then probabilities x V^T with a broadcast A (raster_b path)."""
    torch.manual_seed(1)
    nb, nh, Ntok, d = 3, 8, 256, 40
    Cc = nh * d
    Q = torch.randn(nb, Ntok, Cc, device="cuda").half()
    Km = torch.randn(Ntok, Cc, device="cuda").half()
    S = torch.zeros(nb, nh, Ntok, Ntok, device="cuda")
    gemm16(Q, Km, S, M=Ntok, N_=Ntok, K=d, lda=Cc, sAb=Ntok * Cc, sAh=d, ldb=Cc, sBb=0, sBh=d, ldd=Ntok,
           sDb=nh * Ntok * Ntok, sDh=Ntok * Ntok, nb=nb, nh=nh, alpha=d ** -0.5, out16=False)
    q = Q.double().view(nb, Ntok, nh, d).permute(0, 2, 1, 3)
    k = Km.double().view(Ntok, nh, d).permute(1, 0, 2)
    assert rel(S, torch.einsum("bhid,hjd->bhij", q, k) * d ** -0.5) < 1e-5
    P = torch.softmax(S[0], -1).half()                                   # [nh][N][N], shared by all tangents
    dVt = torch.randn(nb, nh, d, Ntok, device="cuda").half()
    O = torch.zeros(nb, Ntok, Cc, device="cuda", dtype=torch.float16)
    gemm16(P, dVt, O, M=Ntok, N_=d, K=Ntok, lda=Ntok, sAb=0, sAh=Ntok * Ntok, ldb=Ntok, sBb=nh * d * Ntok, sBh=d * Ntok,
           ldd=Cc, sDb=Ntok * Cc, sDh=d, nb=nb, nh=nh)
    ref = torch.einsum("hij,bhdj->bihd", P.double(), dVt.double()).reshape(nb, Ntok, Cc)
    assert rel(O, ref) < 6e-4


@pytest.mark.parametrize("M,Nn,K,conv", [(320, 1280, 1280, 1), (320, 1280, 11520, 0), (64, 1280, 1280, 1), (1280, 640, 640, 1),
                                         (100, 200, 2000, 0)])
def test_gemm_split_k_is_deterministic_and_matches_unsplit(M, Nn, K, conv):
    """Small-M weight-streaming shapes take the split-K path: same result (to accumulation order) as the unsplit
    kernel, bit-identical from run to run, residual + bias + rounding applied once by the reducing CTA."""
    torch.manual_seed(7)
    if conv:
        nb = M // 64
        A = torch.randn(nb, 8, 8, K, device="cuda")
        B = torch.randn(Nn, 9 * K, device="cuda") / math.sqrt(9 * K)
        kw = dict(M=M, N_=Nn, K=K, lda=K, ldb=9 * K, ldd=Nn, nb=nb, conv=1, H=8, W=8)
    else:
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(Nn, K, device="cuda") / math.sqrt(K)
        kw = dict(M=M, N_=Nn, K=K, lda=K, ldb=K, ldd=Nn)
    bias = torch.randn(Nn, device="cuda")
    R = torch.randn(M, Nn, device="cuda")
    outs = []
    for splitk in (True, True, False):
        D = torch.full((M, Nn), float("nan"), device="cuda")
        gemm(A, B, D, bias=bias, R=R, ldr=Nn, alpha=0.5, beta=2.0, rnd=1, splitk=splitk, **kw)
        outs.append(D)
    assert torch.equal(outs[0], outs[1])
    assert rel(outs[0], outs[2]) < 2e-4               # both TF32-rounded on store: one-ulp flips allowed
    assert not torch.isnan(outs[0]).any()


def test_gemm_two_segments_and_inplace_residual():
    torch.manual_seed(0)
    M, Nn, K1, K2 = 512, 320, 640, 320
    A1, A2 = torch.randn(M, K1, device="cuda"), torch.randn(M, K2, device="cuda")
    B = torch.randn(Nn, K1 + K2, device="cuda")
    D = torch.randn(M, Nn, device="cuda")
    D0 = D.clone()
    gemm(A1, B, D, M=M, N_=Nn, K=K1, lda=K1, ldb=K1 + K2, ldd=Nn, R=D, ldr=Nn, beta=1.0,
         seg2=(A2, B[:, K1:], K2, K2, K1 + K2))
    ref = tf32_trunc(torch.cat([A1, A2], 1)).double() @ tf32_trunc(B).double().T + D0.double()
    assert rel(D, ref) < 1e-5


@pytest.mark.parametrize("nb,nh,Ntok,d,Nk", [(3, 8, 256, 40, 256), (2, 8, 64, 160, 64), (5, 2, 100, 32, 77), (2, 5, 1024, 64, 1024)])
def test_gemm_attention_shapes(nb, nh, Ntok, d, Nk):
    """S[b,h] = scale * Q_b[:, h*d:(h+1)*d] K[:, h*d:(h+1)*d]^T with K broadcast over b (cross/self primal K)."""
    torch.manual_seed(1)
    Cc = nh * d
    Q = torch.randn(nb, Ntok, Cc, device="cuda")
    Km = torch.randn(Nk, Cc, device="cuda")
    ld = (Nk + 3) // 4 * 4
    S = torch.zeros(nb, nh, Ntok, ld, device="cuda")
    gemm(Q, Km, S, M=Ntok, N_=Nk, K=d, lda=Cc, sAb=Ntok * Cc, sAh=d, ldb=Cc, sBb=0, sBh=d, ldd=ld,
         sDb=nh * Ntok * ld, sDh=Ntok * ld, nb=nb, nh=nh, alpha=d ** -0.5)
    q = tf32_trunc(Q).double().view(nb, Ntok, nh, d).permute(0, 2, 1, 3)
    k = tf32_trunc(Km).double().view(Nk, nh, d).permute(1, 0, 2)
    ref = torch.einsum("bhid,hjd->bhij", q, k) * d ** -0.5
    assert rel(S[..., :Nk], ref) < 1e-5
    # O[b][:, h*d:(h+1)*d] = P[b,h] Vt[h]^T  with Vt [h][d][Nk]
    Vt = torch.randn(nh, d, ld, device="cuda")
    O = torch.zeros(nb, Ntok, Cc, device="cuda")
    gemm(S, Vt, O, M=Ntok, N_=d, K=Nk, lda=ld, sAb=nh * Ntok * ld, sAh=Ntok * ld, ldb=ld, sBh=d * ld, ldd=Cc,
         sDb=Ntok * Cc, sDh=d, nb=nb, nh=nh)
    ref = torch.einsum("bhij,hdj->bihd", tf32_trunc(S)[..., :Nk].double(), tf32_trunc(Vt)[..., :Nk].double()).reshape(nb, Ntok, Cc)
    assert rel(O, ref) < 1e-5


@pytest.mark.parametrize("nb,H,W,Ci,Co", [(1, 64, 64, 320, 320), (5, 8, 8, 1280, 1280), (3, 16, 16, 64, 96), (2, 32, 32, 640, 320),
                                          (5, 4, 4, 64, 64), (1, 2, 2, 32, 32), (2, 12, 12, 64, 64), (1, 96, 96, 32, 64)])
def test_gemm_conv3x3(nb, H, W, Ci, Co):
    torch.manual_seed(2)
    x = torch.randn(nb, H, W, Ci, device="cuda")                      # NHWC
    w = torch.randn(Co, Ci, 3, 3, device="cuda") / math.sqrt(9 * Ci)
    bias = torch.randn(Co, device="cuda")
    fwd = torch.empty(Co, 9, Ci, device="cuda")
    bwd = torch.empty(Ci, 9, Co, device="cuda")
    _ok(N.leaf("pbk_pack_conv3x3")(_p(w), Co, Ci, _p(fwd), _p(bwd), 0, _st()))
    assert torch.equal(fwd, w.permute(0, 2, 3, 1).reshape(Co, 9, Ci))
    assert torch.equal(bwd, w.flip(2, 3).permute(1, 2, 3, 0).reshape(Ci, 9, Co))
    y = torch.zeros(nb, H, W, Co, device="cuda")
    gemm(x, fwd, y, M=nb * H * W, N_=Co, K=Ci, lda=Ci, ldb=9 * Ci, ldd=Co, nb=nb, conv=1, H=H, W=W, bias=bias)
    ref = F.conv2d(tf32_trunc(x).double().permute(0, 3, 1, 2), tf32_trunc(w).double(), bias.double(), padding=1).permute(0, 2, 3, 1)
    assert rel(y, ref) < 5e-5                          # fp32 accumulation over K = 9 * Ci up to 11520
    # transpose conv (bwd-data) through the flipped pack
    gy = torch.randn(nb, H, W, Co, device="cuda")
    gx = torch.zeros(nb, H, W, Ci, device="cuda")
    gemm(gy, bwd, gx, M=nb * H * W, N_=Ci, K=Co, lda=Co, ldb=9 * Co, ldd=Ci, nb=nb, conv=1, H=H, W=W)
    ref = F.conv_transpose2d(tf32_trunc(gy).double().permute(0, 3, 1, 2), tf32_trunc(w).double(), padding=1).permute(0, 2, 3, 1)
    assert rel(gx, ref) < 5e-5


def test_groupnorm_fwd_and_lin():
    torch.manual_seed(3)
    nb, HW, Cc, G = 3, 256, 320, 32
    x = torch.randn(1, HW, Cc, device="cuda") * 2 + 0.5
    gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    mean, rstd = torch.empty(G, device="cuda"), torch.empty(G, device="cuda")
    nfl = N.raw().pbk_gn_tmp_floats
    nfl.restype = C.c_size_t
    tmp = torch.empty(max(nfl(HW, Cc, G, nb), nfl(HW, Cc, G, 1)), device="cuda")
    _ok(N.leaf("pbk_gn_stats")(_p(x), 1, HW, Cc, G, C.c_float(1e-5), _p(mean), _p(rstd), _p(tmp), _st()))
    y = torch.empty_like(x)
    for silu in (0, 1):
        _ok(N.leaf("pbk_gn_apply")(_p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), 1, HW, Cc, G, silu, 0, _p(y), _st()))

        def f(z):
            o = F.group_norm(z.permute(0, 2, 1), G, gamma, beta, 1e-5).permute(0, 2, 1)
            return F.silu(o) if silu else o
        assert rel(y, f(x)) < 1e-5
        t = torch.randn(nb, HW, Cc, device="cuda")
        out = torch.empty_like(t)
        _ok(N.leaf("pbk_gn_lin")(_p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), HW, Cc, G, silu, _p(t), nb, 0, _p(out),
                                 C.c_float(0), 0, _p(tmp), 0, C.c_long(0), _st()))
        ref = torch.cat([torch.func.jvp(f, (x,), (t[i:i + 1],))[1] for i in range(nb)])
        assert rel(out, ref) < 2e-5
        _ok(N.leaf("pbk_gn_lin")(_p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), HW, Cc, G, silu, _p(t), nb, 1, _p(out),
                                 C.c_float(0), 0, _p(tmp), 0, C.c_long(0), _st()))
        ref = torch.cat([torch.func.vjp(f, x)[1](t[i:i + 1])[0] for i in range(nb)])
        assert rel(out, ref) < 2e-5


@pytest.mark.parametrize("nb,HW,Cc,G", [(5, 4096, 320, 32), (5, 64, 1280, 32), (2, 1024, 160, 32), (3, 256, 96, 32), (1, 4096, 128, 32),
                                       (2, 65536, 128, 32)])
def test_groupnorm_lin_both_paths(nb, HW, Cc, G):
    """pbk_gn_lin through the one-launch group kernel (4 / 2 / 1 channels per access: cpg 40, 10, 5, 3) and through the chunked
    three-launch path (few (image, group) pairs on a large tensor), fp32 / accumulate / fp16 outputs, against torch autograd."""
    torch.manual_seed(5)
    x = torch.randn(1, HW, Cc, device="cuda") * 1.5 + 0.3
    gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    mean, rstd = torch.empty(G, device="cuda"), torch.empty(G, device="cuda")
    nfl = N.raw().pbk_gn_tmp_floats
    nfl.restype = C.c_size_t
    tmp = torch.empty(max(nfl(HW, Cc, G, nb), nfl(HW, Cc, G, 1)), device="cuda")
    _ok(N.leaf("pbk_gn_stats")(_p(x), 1, HW, Cc, G, C.c_float(1e-5), _p(mean), _p(rstd), _p(tmp), _st()))
    f = lambda z: F.silu(F.group_norm(z.permute(0, 2, 1), G, gamma, beta, 1e-5).permute(0, 2, 1))
    t = torch.randn(nb, HW, Cc, device="cuda")
    for mode in (0, 1):
        if mode == 0:
            ref = torch.cat([torch.func.jvp(f, (x,), (t[i:i + 1],))[1] for i in range(nb)])
        else:
            ref = torch.cat([torch.func.vjp(f, x)[1](t[i:i + 1])[0] for i in range(nb)])
        prev = torch.randn_like(t)
        out = prev.clone()
        _ok(N.leaf("pbk_gn_lin")(_p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), HW, Cc, G, 1, _p(t), nb, mode, _p(out),
                                 C.c_float(1.0), 0, _p(tmp), 0, C.c_long(0), _st()))
        assert rel(out, ref + prev) < 3e-5
        if Cc % 8 == 0:
            o16 = torch.zeros(nb, HW, Cc, device="cuda", dtype=torch.float16)
            _ok(N.leaf("pbk_gn_lin")(_p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), HW, Cc, G, 1, _p(t), nb, mode, _p(o16),
                                     C.c_float(0), 2, _p(tmp), 0, C.c_long(0), _st()))
            assert rel(o16.float(), ref) < 1e-3


def test_attn_delta():
    """delta[b][h][i] = sum_c go[b][i][h d + c] o[i][h d + c] (the row term of the softmax cotangent)."""
    torch.manual_seed(8)
    for nb, Ntok, H, d in ((5, 4096, 8, 40), (2, 100, 3, 12), (1, 64, 1, 512)):
        Cc = H * d
        go, o = torch.randn(nb, Ntok, Cc, device="cuda"), torch.randn(Ntok, Cc, device="cuda")
        delta = torch.empty(nb, H, Ntok, device="cuda")
        _ok(N.leaf("pbk_attn_delta")(_p(go), C.c_long(Cc), _p(o), C.c_long(Cc), nb, Ntok, H, d, _p(delta), 0, 0, C.c_long(0), _st()))
        ref = (go.double() * o.double()[None]).view(nb, Ntok, H, d).sum(-1).permute(0, 2, 1)
        assert rel(delta, ref) < 1e-5


def test_layernorm_geglu_softmax():
    torch.manual_seed(4)
    nb, rows, Cc = 3, 200, 320
    x = torch.randn(rows, Cc, device="cuda") + 0.3
    gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    y, mean, rstd = torch.empty_like(x), torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    _ok(N.leaf("pbk_ln_fwd")(_p(x), C.c_long(rows), Cc, _p(gamma), _p(beta), C.c_float(1e-5), _p(y), _p(mean), _p(rstd), 0, _st()))
    f = lambda z: F.layer_norm(z, (Cc,), gamma, beta, 1e-5)
    assert rel(y, f(x)) < 1e-5
    t = torch.randn(nb, rows, Cc, device="cuda")
    out = torch.empty_like(t)
    for mode in (0, 1):
        _ok(N.leaf("pbk_ln_lin")(_p(x), _p(mean), _p(rstd), _p(gamma), C.c_long(rows), Cc, _p(t), nb, mode, _p(out), C.c_float(0), 0, 0, C.c_long(0), _st()))
        if mode == 0:
            ref = torch.stack([torch.func.jvp(f, (x,), (t[i],))[1] for i in range(nb)])
        else:
            ref = torch.stack([torch.func.vjp(f, x)[1](t[i])[0] for i in range(nb)])
        assert rel(out, ref) < 2e-5
    # GEGLU
    Fd = 4 * Cc
    h = torch.randn(rows, 2 * Fd, device="cuda")
    g = lambda z: z[..., :Fd] * F.gelu(z[..., Fd:])
    yy = torch.empty(rows, Fd, device="cuda")
    _ok(N.leaf("pbk_geglu_fwd")(_p(h), C.c_long(rows), Fd, _p(yy), 0, 0, _st()))
    assert rel(yy, g(h)) < 1e-5
    hprep = h.clone()                               # prepared cache [gelu(g) | a gelu'(g)], written in place by the primal kernel
    _ok(N.leaf("pbk_geglu_fwd")(_p(hprep), C.c_long(rows), Fd, _p(yy), 0, 1, _st()))
    assert rel(yy, g(h)) < 1e-5 and rel(hprep[:, :Fd], F.gelu(h[:, Fd:])) < 1e-5
    dh = torch.randn(nb, rows, 2 * Fd, device="cuda")
    dy = torch.empty(nb, rows, Fd, device="cuda")
    _ok(N.leaf("pbk_geglu_jvp")(_p(hprep), C.c_long(rows), _p(dh), nb, Fd, _p(dy), 0, 0, C.c_long(0), _st()))
    assert rel(dy, torch.stack([torch.func.jvp(g, (h,), (dh[i],))[1] for i in range(nb)])) < 2e-5
    gy = torch.randn(nb, rows, Fd, device="cuda")
    gh = torch.empty(nb, rows, 2 * Fd, device="cuda")
    _ok(N.leaf("pbk_geglu_vjp")(_p(hprep), C.c_long(rows), _p(gy), nb, Fd, _p(gh), 0, 0, C.c_long(0), _st()))
    assert rel(gh, torch.stack([torch.func.vjp(g, h)[1](gy[i])[0] for i in range(nb)])) < 2e-5
    # softmax + its linearisation, short (warp) and long (block) rows
    for cols in (77, 256, 4096):
        ld = (cols + 3) // 4 * 4
        S = torch.randn(rows, ld, device="cuda") * 3
        P = S.clone()
        _ok(N.leaf("pbk_softmax_fwd")(_p(P), C.c_long(rows), cols, C.c_long(ld), 0, _st()))
        sm = lambda z: torch.softmax(z, -1)
        assert rel(P[:, :cols], sm(S[:, :cols])) < 1e-5
        dS = torch.randn(nb, rows, ld, device="cuda")
        ref = torch.stack([torch.func.jvp(sm, (S[:, :cols],), (dS[i, :, :cols],))[1] for i in range(nb)])
        _ok(N.leaf("pbk_softmax_lin")(_p(P), C.c_long(rows), _p(dS), nb, cols, C.c_long(ld), 0, 0, C.c_long(0), _st()))
        assert rel(dS[..., :cols], ref) < 2e-5


def test_data_movement():
    torch.manual_seed(5)
    nb, H, W, Cc = 2, 8, 8, 64
    x = torch.randn(nb, H, W, Cc, device="cuda")
    y = torch.empty(nb, 2 * H, 2 * W, Cc, device="cuda")
    _ok(N.leaf("pbk_upsample2x")(_p(x), nb, H, W, Cc, _p(y), 0, _st()))
    ref = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(y, ref)
    gx = torch.empty_like(x)
    _ok(N.leaf("pbk_upsample2x_vjp")(_p(y), nb, H, W, Cc, _p(gx), C.c_float(0), 0, _st()))
    assert rel(gx, 4 * x) < 1e-6
    for pad in (1, 0):                                 # SD pad 1 / DDPM (0,1,0,1) pad
        Ho = Wo = H // 2
        col = torch.empty(nb, Ho, Wo, 9, Cc, device="cuda")
        _ok(N.leaf("pbk_im2col_s2")(_p(x), nb, H, W, Cc, pad, Ho, Wo, _p(col), 0, _st()))
        xp = x.permute(0, 3, 1, 2)
        xp = F.pad(xp, (1, 1, 1, 1)) if pad else F.pad(xp, (0, 1, 0, 1))
        ref = F.unfold(xp, 3, stride=2).view(nb, Cc, 9, Ho, Wo).permute(0, 3, 4, 2, 1)
        assert torch.equal(col, ref.contiguous())
        g = torch.randn_like(col)
        gx = torch.empty_like(x)
        _ok(N.leaf("pbk_col2im_s2")(_p(g), nb, H, W, Cc, pad, Ho, Wo, _p(gx), C.c_float(0), 0, _st()))
        fold = F.fold(g.permute(0, 4, 3, 1, 2).reshape(nb, Cc * 9, Ho * Wo), (H + 2 * pad if pad else H + 1,) * 2, 3, stride=2)
        fold = fold[:, :, 1:-1, 1:-1] if pad else fold[:, :, :-1, :-1]
        assert rel(gx, fold.permute(0, 2, 3, 1)) < 1e-6
    # transpose (V^T per head) and strided copy
    src = torch.randn(3, 100, 48, device="cuda")
    dst = torch.zeros(3, 48, 104, device="cuda")
    _ok(N.leaf("pbk_transpose")(_p(dst), C.c_long(104), C.c_long(48 * 104), C.c_long(0), _p(src), C.c_long(48), C.c_long(100 * 48),
                                C.c_long(0), 3, 1, 100, 48, C.c_float(0), 0, _st()))
    assert torch.equal(dst[:, :, :100], src.transpose(1, 2))
    # thin direct convs (conv_in and its transpose)
    xi = torch.randn(nb, H, W, 4, device="cuda")
    w = torch.randn(32, 4, 3, 3, device="cuda")
    b = torch.randn(32, device="cuda")
    fwd, bwd = torch.empty(32, 9, 4, device="cuda"), torch.empty(4, 9, 32, device="cuda")
    _ok(N.leaf("pbk_pack_conv3x3")(_p(w), 32, 4, _p(fwd), _p(bwd), 0, _st()))
    yo = torch.empty(nb, H, W, 32, device="cuda")
    _ok(N.leaf("pbk_conv3x3_direct")(_p(xi), nb, H, W, 4, _p(fwd), _p(b), 32, _p(yo), C.c_float(0), 0, _st()))
    assert rel(yo, F.conv2d(xi.permute(0, 3, 1, 2), w, b, padding=1).permute(0, 2, 3, 1)) < 1e-5
    go = torch.randn(nb, H, W, 32, device="cuda")
    gi = torch.empty(nb, H, W, 4, device="cuda")
    _ok(N.leaf("pbk_conv3x3_direct")(_p(go), nb, H, W, 32, _p(bwd), None, 4, _p(gi), C.c_float(0), 0, _st()))
    assert rel(gi, F.conv_transpose2d(go.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)) < 1e-5


@pytest.mark.parametrize("nb,H,W,Cin,Cout", [(5, 64, 64, 4, 320), (2, 19, 23, 3, 128), (3, 16, 16, 4, 64), (1, 8, 8, 4, 1280)])
def test_thin_convs_blocked(nb, H, W, Cin, Cout):
    """conv_in (Cin 3 / 4 -> Cout, register-blocked) and its transpose (filter in shared memory), incl. beta accumulation."""
    torch.manual_seed(11)
    xi = torch.randn(nb, H, W, Cin, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / 6
    b = torch.randn(Cout, device="cuda")
    fwd, bwd = torch.empty(Cout, 9, Cin, device="cuda"), torch.empty(Cin, 9, Cout, device="cuda")
    _ok(N.leaf("pbk_pack_conv3x3")(_p(w), Cout, Cin, _p(fwd), _p(bwd), 0, _st()))
    y0 = torch.randn(nb, H, W, Cout, device="cuda")
    yo = y0.clone()
    _ok(N.leaf("pbk_conv3x3_direct")(_p(xi), nb, H, W, Cin, _p(fwd), _p(b), Cout, _p(yo), C.c_float(0.5), 0, _st()))
    ref = F.conv2d(xi.permute(0, 3, 1, 2), w, b, padding=1).permute(0, 2, 3, 1) + 0.5 * y0
    assert rel(yo, ref) < 1e-5
    go = torch.randn(nb, H, W, Cout, device="cuda")
    g0 = torch.randn(nb, H, W, Cin, device="cuda")
    gi = g0.clone()
    _ok(N.leaf("pbk_conv3x3_direct")(_p(go), nb, H, W, Cout, _p(bwd), None, Cin, _p(gi), C.c_float(1.0), 0, _st()))
    ref = F.conv_transpose2d(go.permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1) + g0
    assert rel(gi, ref) < 1e-5


@pytest.mark.parametrize("c2", [0, 1, 2])
@pytest.mark.parametrize("Mr,Nc,d,nb,nh,nseg,mode", [(256, 256, 40, 2, 2, 2, 0), (384, 128, 64, 2, 2, 1, 2), (1024, 1024, 40, 3, 8, 2, 0),
                                                       (128, 640, 32, 2, 1, 2, 1), (4096, 4096, 40, 1, 2, 2, 0), (300, 104, 48, 2, 2, 2, 2),
                                                       (512, 512, 56, 2, 1, 2, 1), (2304, 2304, 64, 1, 5, 2, 0), (200, 328, 16, 1, 3, 1, 1)])
def test_attn_lin_fused_fp16_operands(Mr, Nc, d, nb, nh, nseg, mode, c2):
    """PbAttnLin with p16 = 1: Pm (pre-scaled by p_scale = Nc), C1 and C2 in fp16, 64-column steps, kind::f16 accumulating products."""
    torch.manual_seed(7)
    Cc = nh * d
    ldp = (Nc + 7) // 8 * 8
    A0 = torch.randn(nb, Mr, Cc, device="cuda"); B0 = torch.randn(Nc, Cc, device="cuda")
    A1 = torch.randn(Mr, Cc, device="cuda"); B1 = torch.randn(nb, Nc, Cc, device="cuda")
    Pm = torch.softmax(torch.randn(nh, Mr, ldp, device="cuda") * 2, -1).contiguous()
    Pm[..., Nc:] = 0
    P16 = (Pm * Nc).half().contiguous()
    C1 = torch.randn(nh, d, ldp, device="cuda").half()
    C2 = torch.randn(nb, nh, d, ldp, device="cuda").half()
    O = torch.randn(Mr, Cc, device="cuda"); R = torch.randn(nb, Mr, Cc, device="cuda")
    delta = torch.randn(nb, nh, Mr if mode == 1 else Nc, device="cuda") if mode else None
    D = R.clone()
    D2 = torch.full((nb, Mr, Cc), float("nan"), device="cuda")
    a = N.PbAttnLin()
    a.Mr, a.Nc, a.d, a.nb, a.nh, a.nseg = Mr, Nc, d, nb, nh, nseg
    s0 = a.seg[0]
    s0.A, s0.lda, s0.sAb, s0.sAh, s0.B, s0.ldb, s0.sBb, s0.sBh = A0.data_ptr(), Cc, Mr * Cc, d, B0.data_ptr(), Cc, 0, d
    s1 = a.seg[1]
    s1.A, s1.lda, s1.sAb, s1.sAh, s1.B, s1.ldb, s1.sBb, s1.sBh = A1.data_ptr(), Cc, 0, d, B1.data_ptr(), Cc, Nc * Cc, d
    a.alpha1, a.alpha2, a.beta = d ** -0.5, 0.7, 1.0
    a.Pm, a.ldp, a.sPh = P16.data_ptr(), ldp, Mr * ldp
    a.delta, a.delta_mode = (delta.data_ptr() if mode else None), mode
    a.want_rsum, a.O, a.ldo = int(mode == 0), O.data_ptr(), Cc
    a.C1, a.ldc, a.sCh = C1.data_ptr(), ldp, d * ldp
    a.D, a.ldd, a.sDb, a.R, a.ldr, a.sRb, a.round_tf32 = D.data_ptr(), Cc, Mr * Cc, D.data_ptr(), Cc, Mr * Cc, 0
    a.p16, a.p_scale = 1, float(Nc)
    if c2:
        a.C2, a.ldc2, a.sC2h, a.sC2b = C2.data_ptr(), ldp, d * ldp, nh * d * ldp
    if c2 == 2:
        a.D2, a.ldd2, a.sD2b = D2.data_ptr(), Cc, Mr * Cc
    _ok(N.leaf("pbk_attn_lin")(C.byref(a), _st()))
    t = lambda z: tf32_trunc(z).double()
    S = torch.einsum("bihd,jhd->bhij", t(A0).view(nb, Mr, nh, d), t(B0).view(Nc, nh, d))
    if nseg == 2:
        S = S + torch.einsum("ihd,bjhd->bhij", t(A1).view(Mr, nh, d), t(B1).view(nb, Nc, nh, d))
    S = S * d ** -0.5
    if mode == 1:
        S = S - delta.double()[..., :, None]
    if mode == 2:
        S = S - delta.double()[..., None, :]
    Ps = P16.double()[None, :, :, :Nc]                                 # the scaled probabilities the kernel reads
    Tr = (Ps * S).float().half().double()                              # T is rounded to fp16 before the second contraction
    acc = torch.einsum("bhij,hnj->bihn", Tr, C1.double()[..., :Nc]).reshape(nb, Mr, Cc) / Nc
    if c2:
        e2 = torch.einsum("hij,bhnj->bihn", P16.double()[..., :Nc], C2.double()[..., :Nc]).reshape(nb, Mr, Cc) / Nc
        if c2 == 1:
            acc = acc + e2
        else:
            assert rel(D2, e2) < 1e-5, rel(D2, e2)
    ref = 0.7 * acc + R.double()
    if mode == 0:
        rs = Tr.sum(-1) / Nc
        ref = ref - (rs.permute(0, 2, 1)[..., None] * O.double().view(Mr, nh, d)[None]).reshape(nb, Mr, Cc)
    assert rel(D, ref) < 3e-4, rel(D, ref)                             # a few T elements round the other way (fp32 vs fp64 S)


def test_attn_lin_fp16_range_peaked_and_uniform_rows():
    """fp16 probability path at the two ends of the range: near one-hot rows with large scores (trained attention) and uniform rows
    with tiny scores, p_scale = sqrt(Nc) as the engine chooses it: finite and accurate."""
    torch.manual_seed(9)
    Mr = Nc = 1024; d = 40; nb = 2; nh = 2; Cc = nh * d
    for sharp, amp in ((40.0, 30.0), (0.0, 1e-2)):
        A0 = torch.randn(nb, Mr, Cc, device="cuda") * amp; B0 = torch.randn(Nc, Cc, device="cuda")
        Pm = torch.softmax(torch.randn(nh, Mr, Nc, device="cuda") * sharp, -1).contiguous()
        ps = 2.0 ** round(math.log2(math.sqrt(Nc)))
        P16 = (Pm * ps).half().contiguous()
        C1 = torch.randn(nh, d, Nc, device="cuda").half()
        O = torch.randn(Mr, Cc, device="cuda")
        D = torch.zeros(nb, Mr, Cc, device="cuda")
        a = N.PbAttnLin()
        a.Mr, a.Nc, a.d, a.nb, a.nh, a.nseg = Mr, Nc, d, nb, nh, 1
        s0 = a.seg[0]
        s0.A, s0.lda, s0.sAb, s0.sAh, s0.B, s0.ldb, s0.sBb, s0.sBh = A0.data_ptr(), Cc, Mr * Cc, d, B0.data_ptr(), Cc, 0, d
        a.alpha1, a.alpha2, a.beta = d ** -0.5, 1.0, 0.0
        a.Pm, a.ldp, a.sPh = P16.data_ptr(), Nc, Mr * Nc
        a.want_rsum, a.O, a.ldo = 1, O.data_ptr(), Cc
        a.C1, a.ldc, a.sCh = C1.data_ptr(), Nc, d * Nc
        a.D, a.ldd, a.sDb, a.round_tf32 = D.data_ptr(), Cc, Mr * Cc, 0
        a.p16, a.p_scale = 1, ps
        _ok(N.leaf("pbk_attn_lin")(C.byref(a), _st()))
        assert torch.isfinite(D).all()
        S = torch.einsum("bihd,jhd->bhij", tf32_trunc(A0).double().view(nb, Mr, nh, d), tf32_trunc(B0).double().view(Nc, nh, d)) * d ** -0.5
        T = Pm.double()[None] * S
        ref = torch.einsum("bhij,hnj->bihn", T, C1.double()).reshape(nb, Mr, Cc)
        ref = ref - (T.sum(-1).permute(0, 2, 1)[..., None] * O.double().view(Mr, nh, d)[None]).reshape(nb, Mr, Cc)
        assert rel(D, ref) < 2e-3, (sharp, amp, rel(D, ref))


def test_orthonormalize_matches_svd():
    torch.manual_seed(6)
    for k, n in ((5, 16384), (16, 16384), (2, 196608), (50, 4096)):
        Wm = torch.randn(k, n, device="cuda") * torch.linspace(3, 0.5, k, device="cuda")[:, None]
        Vp = torch.linalg.svd(Wm + 0.01 * torch.randn_like(Wm), full_matrices=False)[2].contiguous()
        G, M = torch.empty(k, k, dtype=torch.float64, device="cuda"), torch.empty(k, k, dtype=torch.float64, device="cuda")
        Rm, sv, V, met = (torch.empty(k, k, device="cuda"), torch.empty(k, device="cuda"), torch.empty(k, n, device="cuda"),
                          torch.empty(2, device="cuda"))
        _ok(N.leaf("pbk_gram2")(_p(Wm), _p(Vp), k, C.c_long(n), _p(G), _p(M), _st()))
        _ok(N.leaf("pbk_jacobi")(_p(G), _p(M), k, _p(Rm), _p(sv), _st()))
        _ok(N.leaf("pbk_rotate")(_p(Wm), _p(Rm), _p(Vp), k, C.c_long(n), C.c_float(1e-4), C.c_float(1e-5), _p(V), _p(met), _st()))
        _, s, Vh = torch.linalg.svd(Wm.double(), full_matrices=False)
        assert torch.allclose(sv.double(), s.sqrt(), rtol=1e-5)
        cos = (V.double() * Vh).sum(1).abs()
        assert float(cos.min()) > 1 - 1e-6
        assert torch.allclose(V @ V.T, torch.eye(k, device="cuda"), atol=1e-5)
        assert float((V * Vp).sum(1).min()) > 0          # sign continuity with Vprev
        assert abs(float(met[0]) - float((V - Vp).pow(2).sum())) < 1e-3 * float((V - Vp).pow(2).sum()) + 1e-6


@pytest.mark.parametrize("c2", [0, 1, 2])          # no second product / folded into D / separate output D2
@pytest.mark.parametrize("Mr,Nc,d,nb,nh,nseg,mode", [(256, 256, 40, 2, 2, 2, 0), (200, 333, 16, 1, 3, 1, 1), (384, 128, 64, 2, 2, 1, 2),
                                                       (1024, 1024, 40, 3, 8, 2, 0), (128, 640, 32, 2, 1, 2, 1), (4096, 4096, 40, 1, 2, 2, 0),
                                                       (300, 97, 48, 2, 2, 2, 2), (160, 2048, 8, 1, 2, 1, 0), (512, 512, 56, 2, 1, 2, 1),
                                                       (256, 544, 36, 2, 2, 2, 0), (2304, 2304, 64, 2, 5, 2, 0),
                                                       (1024, 1024, 80, 2, 8, 2, 0), (512, 512, 96, 1, 2, 1, 2), (256, 384, 72, 2, 2, 2, 1)])
def test_attn_lin_fused(Mr, Nc, d, nb, nh, nseg, mode, c2):
    """D = alpha2 * ([Pm o (alpha1 * sum_seg A B^T - delta)] C1^T [+ Pm C2^T]) - rowsum(.) o O + beta * R ; D2 = Pm C2^T   (PbAttnLin)."""
    torch.manual_seed(7)
    Cc = nh * d
    ldp = (Nc + 3) // 4 * 4
    A0 = torch.randn(nb, Mr, Cc, device="cuda"); B0 = torch.randn(Nc, Cc, device="cuda")          # tangent x primal
    A1 = torch.randn(Mr, Cc, device="cuda"); B1 = torch.randn(nb, Nc, Cc, device="cuda")          # primal x tangent
    Pm = torch.softmax(torch.randn(nh, Mr, ldp, device="cuda") * 2, -1).contiguous()
    C1 = torch.randn(nh, d, ldp, device="cuda")
    O = torch.randn(Mr, Cc, device="cuda"); R = torch.randn(nb, Mr, Cc, device="cuda")
    delta = torch.randn(nb, nh, Mr if mode == 1 else Nc, device="cuda") if mode else None
    D = R.clone()
    a = N.PbAttnLin()
    a.Mr, a.Nc, a.d, a.nb, a.nh, a.nseg = Mr, Nc, d, nb, nh, nseg
    s0 = a.seg[0]
    s0.A, s0.lda, s0.sAb, s0.sAh, s0.B, s0.ldb, s0.sBb, s0.sBh = A0.data_ptr(), Cc, Mr * Cc, d, B0.data_ptr(), Cc, 0, d
    s1 = a.seg[1]
    s1.A, s1.lda, s1.sAb, s1.sAh, s1.B, s1.ldb, s1.sBb, s1.sBh = A1.data_ptr(), Cc, 0, d, B1.data_ptr(), Cc, Nc * Cc, d
    a.alpha1, a.alpha2, a.beta = d ** -0.5, 0.7, 1.0
    a.Pm, a.ldp, a.sPh = Pm.data_ptr(), ldp, Mr * ldp
    a.delta, a.delta_mode = (delta.data_ptr() if mode else None), mode
    a.want_rsum, a.O, a.ldo = int(mode == 0), O.data_ptr(), Cc
    a.C1, a.ldc, a.sCh = C1.data_ptr(), ldp, d * ldp
    a.D, a.ldd, a.sDb, a.R, a.ldr, a.sRb, a.round_tf32 = D.data_ptr(), Cc, Mr * Cc, D.data_ptr(), Cc, Mr * Cc, 0
    C2 = torch.randn(nb, nh, d, ldp, device="cuda")
    D2 = torch.full((nb, Mr, Cc), float("nan"), device="cuda")
    if c2:
        a.C2, a.ldc2, a.sC2h, a.sC2b = C2.data_ptr(), ldp, d * ldp, nh * d * ldp
    if c2 == 2:
        a.D2, a.ldd2, a.sD2b = D2.data_ptr(), Cc, Mr * Cc
    _ok(N.leaf("pbk_attn_lin")(C.byref(a), _st()))
    t = lambda z: tf32_trunc(z).double()
    S = torch.einsum("bihd,jhd->bhij", t(A0).view(nb, Mr, nh, d), t(B0).view(Nc, nh, d))
    if nseg == 2:
        S = S + torch.einsum("ihd,bjhd->bhij", t(A1).view(Mr, nh, d), t(B1).view(nb, Nc, nh, d))
    S = S * d ** -0.5
    if mode == 1:
        S = S - delta.double()[..., :, None]
    if mode == 2:
        S = S - delta.double()[..., None, :]
    T = Pm.double()[None, :, :, :Nc] * S
    # the kernel RNA-rounds T to TF32 before the second contraction
    Tr = T.float()
    Tr = ((Tr.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32).double()
    acc = torch.einsum("bhij,hnj->bihn", Tr, t(C1)[..., :Nc]).reshape(nb, Mr, Cc)
    if c2:
        e2 = torch.einsum("hij,bhnj->bihn", t(Pm)[..., :Nc], t(C2)[..., :Nc]).reshape(nb, Mr, Cc)
        if c2 == 1:
            acc = acc + e2
        else:
            assert rel(D2, e2) < 1e-5, rel(D2, e2)
    ref = 0.7 * acc + R.double()
    if mode == 0:
        rs = Tr.sum(-1)                                               # [nb, nh, Mr]
        ref = ref - (rs.permute(0, 2, 1)[..., None] * O.double().view(Mr, nh, d)[None]).reshape(nb, Mr, Cc)
    assert rel(D, ref) < 2e-4, rel(D, ref)                            # a few T elements round the other way (fp32 vs fp64 S)


# ---- the all-fp16 tangent plan: kernels that read and write halves (io = PB_IN_F16 | PB_OUT_F16 = 6) ----
IO16 = 6


@pytest.mark.parametrize("nb,HW,Cc,G", [(5, 4096, 320, 32), (5, 64, 1280, 32), (2, 1024, 640, 32), (3, 256, 2560, 32), (2, 100, 64, 32),
                                       (1, 4096, 32, 32), (3, 1024, 1920, 32), (2, 256, 960, 32)])
def test_fp16_groupnorm_lin(nb, HW, Cc, G):
    """pbk_gn_lin over fp16 tangents (gn16_sums_k + gn16_apply_k): JVP and VJP, with and without accumulation, against torch
    autograd on the same half-rounded tangent; result correct to fp16 rounding of the output."""
    torch.manual_seed(5)
    x = torch.randn(1, HW, Cc, device="cuda") * 1.5 + 0.3
    gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    mean, rstd = torch.empty(G, device="cuda"), torch.empty(G, device="cuda")
    nfl = N.raw().pbk_gn_tmp_floats
    nfl.restype = C.c_size_t
    tmp = torch.empty(max(nfl(HW, Cc, G, nb), nfl(HW, Cc, G, 1)), device="cuda")
    _ok(N.leaf("pbk_gn_stats")(_p(x), 1, HW, Cc, G, C.c_float(1e-5), _p(mean), _p(rstd), _p(tmp), _st()))
    for silu in (1, 0):
        def f(z):
            o = F.group_norm(z.permute(0, 2, 1), G, gamma, beta, 1e-5).permute(0, 2, 1)
            return F.silu(o) if silu else o
        t16 = torch.randn(nb, HW, Cc, device="cuda").half()
        t = t16.float()
        for mode in (0, 1):
            if mode == 0:
                ref = torch.cat([torch.func.jvp(f, (x,), (t[i:i + 1],))[1] for i in range(nb)])
            else:
                ref = torch.cat([torch.func.vjp(f, x)[1](t[i:i + 1])[0] for i in range(nb)])
            out = torch.full((nb, HW, Cc), float("nan"), device="cuda", dtype=torch.float16)
            _ok(N.leaf("pbk_gn_lin")(_p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), HW, Cc, G, silu, _p(t16), nb, mode, _p(out),
                                     C.c_float(0), IO16, _p(tmp), 0, C.c_long(0), _st()))
            assert rel(out.float(), ref) < 6e-4, (silu, mode, rel(out.float(), ref))
            prev = torch.randn(nb, HW, Cc, device="cuda").half()
            out = prev.clone()
            _ok(N.leaf("pbk_gn_lin")(_p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), HW, Cc, G, silu, _p(t16), nb, mode, _p(out),
                                     C.c_float(1.0), IO16, _p(tmp), 0, C.c_long(0), _st()))
            assert rel(out.float(), ref + prev.float()) < 6e-4


@pytest.mark.parametrize("rows,Cc", [(200, 320), (77, 640), (64, 1280), (130, 64), (50, 2560), (33, 96)])
def test_fp16_layernorm_geglu(rows, Cc):
    torch.manual_seed(4)
    nb = 3
    x = torch.randn(rows, Cc, device="cuda") + 0.3
    gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    y, mean, rstd = torch.empty_like(x), torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    _ok(N.leaf("pbk_ln_fwd")(_p(x), C.c_long(rows), Cc, _p(gamma), _p(beta), C.c_float(1e-5), _p(y), _p(mean), _p(rstd), 0, _st()))
    f = lambda z: F.layer_norm(z, (Cc,), gamma, beta, 1e-5)
    t16 = torch.randn(nb, rows, Cc, device="cuda").half()
    for mode in (0, 1):
        if mode == 0:
            ref = torch.stack([torch.func.jvp(f, (x,), (t16[i].float(),))[1] for i in range(nb)])
        else:
            ref = torch.stack([torch.func.vjp(f, x)[1](t16[i].float())[0] for i in range(nb)])
        prev = torch.randn(nb, rows, Cc, device="cuda").half()
        for acc in (0.0, 1.0):
            out = prev.clone()
            _ok(N.leaf("pbk_ln_lin")(_p(x), _p(mean), _p(rstd), _p(gamma), C.c_long(rows), Cc, _p(t16), nb, mode, _p(out), C.c_float(acc),
                                     IO16, 0, C.c_long(0), _st()))
            assert rel(out.float(), ref + acc * prev.float()) < 6e-4
    Fd = 4 * Cc
    h = torch.randn(rows, 2 * Fd, device="cuda")
    g = lambda z: z[..., :Fd] * F.gelu(z[..., Fd:])
    hprep, ytmp = h.clone(), torch.empty(rows, Fd, device="cuda")
    _ok(N.leaf("pbk_geglu_fwd")(_p(hprep), C.c_long(rows), Fd, _p(ytmp), 0, 1, _st()))
    dh = torch.randn(nb, rows, 2 * Fd, device="cuda").half()
    dy = torch.empty(nb, rows, Fd, device="cuda", dtype=torch.float16)
    _ok(N.leaf("pbk_geglu_jvp")(_p(hprep), C.c_long(rows), _p(dh), nb, Fd, _p(dy), IO16, 0, C.c_long(0), _st()))
    assert rel(dy.float(), torch.stack([torch.func.jvp(g, (h,), (dh[i].float(),))[1] for i in range(nb)])) < 6e-4
    gy = torch.randn(nb, rows, Fd, device="cuda").half()
    gh = torch.empty(nb, rows, 2 * Fd, device="cuda", dtype=torch.float16)
    _ok(N.leaf("pbk_geglu_vjp")(_p(hprep), C.c_long(rows), _p(gy), nb, Fd, _p(gh), IO16, 0, C.c_long(0), _st()))
    assert rel(gh.float(), torch.stack([torch.func.vjp(g, h)[1](gy[i].float())[0] for i in range(nb)])) < 6e-4


def test_fp16_data_movement():
    torch.manual_seed(5)
    nb, H, W, Cc = 2, 8, 8, 64
    x = torch.randn(nb, H, W, Cc, device="cuda").half()
    y = torch.empty(nb, 2 * H, 2 * W, Cc, device="cuda", dtype=torch.float16)
    _ok(N.leaf("pbk_upsample2x")(_p(x), nb, H, W, Cc, _p(y), IO16, _st()))
    ref = F.interpolate(x.float().permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(y.float(), ref)
    g0 = torch.randn(nb, H, W, Cc, device="cuda").half()
    gx = g0.clone()
    yy = torch.randn_like(y)
    _ok(N.leaf("pbk_upsample2x_vjp")(_p(yy), nb, H, W, Cc, _p(gx), C.c_float(1.0), IO16, _st()))
    refv = yy.float().view(nb, H, 2, W, 2, Cc).sum((2, 4)) + g0.float()
    assert rel(gx.float(), refv) < 6e-4
    for pad in (1, 0):
        Ho = Wo = H // 2
        col = torch.empty(nb, Ho, Wo, 9, Cc, device="cuda", dtype=torch.float16)
        _ok(N.leaf("pbk_im2col_s2")(_p(x), nb, H, W, Cc, pad, Ho, Wo, _p(col), IO16, _st()))
        xp = x.float().permute(0, 3, 1, 2)
        xp = F.pad(xp, (1, 1, 1, 1)) if pad else F.pad(xp, (0, 1, 0, 1))
        ref = F.unfold(xp, 3, stride=2).view(nb, Cc, 9, Ho, Wo).permute(0, 3, 4, 2, 1)
        assert torch.equal(col.float(), ref.contiguous())
        g = torch.randn(nb, Ho, Wo, 9, Cc, device="cuda").half()
        gx = torch.empty(nb, H, W, Cc, device="cuda", dtype=torch.float16)
        _ok(N.leaf("pbk_col2im_s2")(_p(g), nb, H, W, Cc, pad, Ho, Wo, _p(gx), C.c_float(0), IO16, _st()))
        fold = F.fold(g.float().permute(0, 4, 3, 1, 2).reshape(nb, Cc * 9, Ho * Wo), (H + 2 * pad if pad else H + 1,) * 2, 3, stride=2)
        fold = fold[:, :, 1:-1, 1:-1] if pad else fold[:, :, :-1, :-1]
        assert rel(gx.float(), fold.permute(0, 2, 3, 1)) < 6e-4
    # accumulating channel-slice copy (concat split / residual fan-in) over halves
    src = torch.randn(50, 96, device="cuda").half()
    d0 = torch.randn(50, 64, device="cuda").half()
    dst = d0.clone()
    _ok(N.leaf("pbk_copy2d")(_p(dst), C.c_long(64), C.c_void_p(src.data_ptr() + 32 * 2), C.c_long(96), C.c_long(50), 64, C.c_float(1.0), IO16, _st()))
    assert rel(dst.float(), d0.float() + src[:, 32:].float()) < 6e-4
    # transposes at the fp32 <-> fp16 ends and between halves (dV^T per head)
    s32 = torch.randn(3, 100, 48, device="cuda")
    for io, sdt, ddt in ((2, torch.float32, torch.float16), (4, torch.float16, torch.float32), (6, torch.float16, torch.float16)):
        src = s32.to(sdt)
        dst = torch.zeros(3, 48, 104, device="cuda", dtype=ddt)
        _ok(N.leaf("pbk_transpose")(_p(dst), C.c_long(104), C.c_long(48 * 104), C.c_long(0), _p(src), C.c_long(48), C.c_long(100 * 48),
                                    C.c_long(0), 3, 1, 100, 48, C.c_float(0), io, _st()))
        assert torch.equal(dst[:, :, :100].float(), src.transpose(1, 2).to(ddt).float())
    # per-head transposes between halves, the engine's layouts (dV^T out of the fused qkv tangent, Obar^T): the 64-row tile kernel
    for nb_, nh_, R_, d_, lds_ in ((3, 4, 1000, 40, 3 * 160), (2, 5, 4096, 64, 320), (2, 2, 300, 80, 160)):
        src = torch.randn(nb_, R_, lds_, device="cuda").half()
        ldd_ = (R_ + 7) // 8 * 8 + 8
        dst = torch.zeros(nb_, nh_, d_, ldd_, device="cuda", dtype=torch.float16)
        off = lds_ - nh_ * d_                                   # heads start at a column offset (the v third of qkv)
        _ok(N.leaf("pbk_transpose")(_p(dst), C.c_long(ldd_), C.c_long(nh_ * d_ * ldd_), C.c_long(d_ * ldd_),
                                    C.c_void_p(src.data_ptr() + 2 * off), C.c_long(lds_), C.c_long(R_ * lds_), C.c_long(d_), nb_, nh_, R_, d_,
                                    C.c_float(0), 6, _st()))
        ref = src[:, :, off:].view(nb_, R_, nh_, d_).permute(0, 2, 3, 1)
        assert torch.equal(dst[..., :R_], ref) and (dst[..., R_:] == 0).all()
    # fp16 -> fp32 staging copy
    hsrc = torch.randn(4096, device="cuda").half()
    f32 = torch.empty(4096, device="cuda")
    _ok(N.leaf("pbk_to_f32")(_p(f32), _p(hsrc), C.c_size_t(4096), _st()))
    assert torch.equal(f32, hsrc.float())
    # attn_delta over an fp16 cotangent
    nbd, Ntok, Hh, d = 3, 200, 4, 40
    go = torch.randn(nbd, Ntok, Hh * d, device="cuda").half(); o = torch.randn(Ntok, Hh * d, device="cuda")
    delta = torch.empty(nbd, Hh, Ntok, device="cuda")
    _ok(N.leaf("pbk_attn_delta")(_p(go), C.c_long(Hh * d), _p(o), C.c_long(Hh * d), nbd, Ntok, Hh, d, _p(delta), 4, 0, C.c_long(0), _st()))
    assert rel(delta, (go.double() * o.double()[None]).view(nbd, Ntok, Hh, d).sum(-1).permute(0, 2, 1)) < 1e-5


@pytest.mark.parametrize("nb,H,W,Cin,Cout", [(5, 64, 64, 4, 320), (2, 19, 23, 3, 128), (3, 16, 16, 4, 32), (2, 8, 8, 4, 1280)])
def test_fp16_thin_convs(nb, H, W, Cin, Cout):
    """conv_in in the all-fp16 plan: fp32 x_t-shaped tangent in -> fp16 out (io 2); its transpose fp16 in -> fp32 out (io 4)."""
    torch.manual_seed(11)
    xi = torch.randn(nb, H, W, Cin, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / 6
    fwd, bwd = torch.empty(Cout, 9, Cin, device="cuda"), torch.empty(Cin, 9, Cout, device="cuda")
    _ok(N.leaf("pbk_pack_conv3x3")(_p(w), Cout, Cin, _p(fwd), _p(bwd), 0, _st()))
    yo = torch.empty(nb, H, W, Cout, device="cuda", dtype=torch.float16)
    _ok(N.leaf("pbk_conv3x3_direct")(_p(xi), nb, H, W, Cin, _p(fwd), None, Cout, _p(yo), C.c_float(0), 2, _st()))
    assert rel(yo.float(), F.conv2d(xi.permute(0, 3, 1, 2), w, None, padding=1).permute(0, 2, 3, 1)) < 6e-4
    go = torch.randn(nb, H, W, Cout, device="cuda").half()
    gi = torch.empty(nb, H, W, Cin, device="cuda")
    _ok(N.leaf("pbk_conv3x3_direct")(_p(go), nb, H, W, Cout, _p(bwd), None, Cin, _p(gi), C.c_float(0), 4, _st()))
    assert rel(gi, F.conv_transpose2d(go.float().permute(0, 3, 1, 2), w, padding=1).permute(0, 2, 3, 1)) < 1e-5


def test_gemm_tf32_operands_fp16_output():
    """fp32 (TF32) operands with an fp16 result: the last products of the materialised attention path in the all-fp16 plan."""
    torch.manual_seed(2)
    M, Nn, K = 300, 200, 96
    A, B = torch.randn(M, K, device="cuda"), torch.randn(Nn, K, device="cuda")
    R = torch.randn(M, Nn, device="cuda").half()
    D = torch.full((M, Nn), float("nan"), device="cuda", dtype=torch.float16)
    gemm(A, B, D, M=M, N_=Nn, K=K, lda=K, ldb=K, ldd=Nn, R=R, ldr=Nn, alpha=0.5, beta=1.0, ab_dtype=0, d_dtype=1)
    ref = 0.5 * (tf32_trunc(A).double() @ tf32_trunc(B).double().T) + R.double()
    assert rel(D, ref) < 6e-4


@pytest.mark.parametrize("case", ["jvp", "vjp_a", "vjp_b", "cross"])
@pytest.mark.parametrize("Mr,Nc,d,nb,nh", [(256, 256, 40, 2, 2), (1024, 1024, 80, 2, 3), (384, 512, 64, 2, 2), (4096, 4096, 40, 1, 2),
                                           (300, 77, 40, 3, 2), (520, 264, 16, 2, 1), (256, 128, 96, 1, 2), (128, 192, 48, 2, 2)])
def test_attn_lin_all_fp16(Mr, Nc, d, nb, nh, case):
    """PbAttnLin with p16 = s16 = 1 (the all-fp16 tangent plan): fp16 score operands (kind::f16 over the head dim, incl. the
    3-of-4-MMA last block of head dim 40 and the 32-byte tail of head dim 80), scaled fp16 probabilities, fp16 C1 / C2, fp16
    outputs; in the roles the engine uses (JVP: two segments + folded P.C2 + row-sum term; VJP-A: row deltas; VJP-B: column
    deltas + separate D2; cross-attention: one segment + row sums) against fp64 on the same half-rounded operands."""
    torch.manual_seed(7)
    Cc = nh * d
    ldp = (Nc + 7) // 8 * 8
    scale = 2.0 ** round(0.5 * math.log2(Nc))
    hf = lambda *s: torch.randn(*s, device="cuda").half()
    A0, B0 = hf(nb, Mr, Cc), hf(Nc, Cc)
    A1, B1 = hf(Mr, Cc), hf(nb, Nc, Cc)
    Pm = torch.softmax(torch.randn(nh, Mr, ldp, device="cuda") * 2, -1).contiguous()
    Pm[..., Nc:] = 0
    P16 = (Pm * scale).half().contiguous()
    C1 = hf(nh, d, ldp)
    C2 = hf(nb, nh, d, ldp)
    O = torch.randn(Mr, Cc, device="cuda")
    nseg = 2 if case == "jvp" else 1
    mode = {"jvp": 0, "cross": 0, "vjp_a": 1, "vjp_b": 2}[case]
    c2 = {"jvp": 1, "cross": 0, "vjp_a": 0, "vjp_b": 2}[case]
    delta = torch.randn(nb, nh, Mr if mode == 1 else Nc, device="cuda") if mode else None
    D = torch.full((nb, Mr, Cc), float("nan"), device="cuda", dtype=torch.float16)
    D2 = torch.full((nb, Mr, Cc), float("nan"), device="cuda", dtype=torch.float16)
    a = N.PbAttnLin()
    a.Mr, a.Nc, a.d, a.nb, a.nh, a.nseg = Mr, Nc, d, nb, nh, nseg
    s0 = a.seg[0]
    s0.A, s0.lda, s0.sAb, s0.sAh, s0.B, s0.ldb, s0.sBb, s0.sBh = A0.data_ptr(), Cc, Mr * Cc, d, B0.data_ptr(), Cc, 0, d
    s1 = a.seg[1]
    s1.A, s1.lda, s1.sAb, s1.sAh, s1.B, s1.ldb, s1.sBb, s1.sBh = A1.data_ptr(), Cc, 0, d, B1.data_ptr(), Cc, Nc * Cc, d
    a.alpha1, a.alpha2, a.beta = d ** -0.5, 0.7, 0.0
    a.Pm, a.ldp, a.sPh = P16.data_ptr(), ldp, Mr * ldp
    a.delta, a.delta_mode = (delta.data_ptr() if mode else None), mode
    a.want_rsum, a.O, a.ldo = int(mode == 0), O.data_ptr(), Cc
    a.C1, a.ldc, a.sCh = C1.data_ptr(), ldp, d * ldp
    a.D, a.ldd, a.sDb, a.round_tf32 = D.data_ptr(), Cc, Mr * Cc, 1
    a.p16, a.p_scale, a.s16 = 1, scale, 1
    if c2:
        a.C2, a.ldc2, a.sC2h, a.sC2b = C2.data_ptr(), ldp, d * ldp, nh * d * ldp
    if c2 == 2:
        a.D2, a.ldd2, a.sD2b = D2.data_ptr(), Cc, Mr * Cc
    _ok(N.leaf("pbk_attn_lin")(C.byref(a), _st()))
    S = torch.einsum("bihd,jhd->bhij", A0.double().view(nb, Mr, nh, d), B0.double().view(Nc, nh, d))
    if nseg == 2:
        S = S + torch.einsum("ihd,bjhd->bhij", A1.double().view(Mr, nh, d), B1.double().view(nb, Nc, nh, d))
    S = S * d ** -0.5
    if mode == 1:
        S = S - delta.double()[..., :, None]
    if mode == 2:
        S = S - delta.double()[..., None, :]
    Ps = P16.double()[None, :, :, :Nc]
    Tr = (Ps * S).float().half().double()
    acc = torch.einsum("bhij,hnj->bihn", Tr, C1.double()[..., :Nc]).reshape(nb, Mr, Cc) / scale
    if c2:
        e2 = torch.einsum("hij,bhnj->bihn", P16.double()[..., :Nc], C2.double()[..., :Nc]).reshape(nb, Mr, Cc) / scale
        if c2 == 1:
            acc = acc + e2
        else:
            assert rel(D2.float(), e2) < 6e-4, rel(D2.float(), e2)
    ref = 0.7 * acc
    if mode == 0:
        rs = Tr.sum(-1) / scale
        ref = ref - (rs.permute(0, 2, 1)[..., None] * O.double().view(Mr, nh, d)[None]).reshape(nb, Mr, Cc)
    assert rel(D.float(), ref) < 8e-4, rel(D.float(), ref)


def test_fp16_kernels_problem_slots():
    """k_slot / p_stride: image b reads the primal tensors of problem b / k_slot (one primal cache per problem at a uniform
    stride): one launch over two problems == two launches over one problem each, bitwise, for the GroupNorm / LayerNorm / GEGLU
    linearisations, attn_delta and the fused attention kernel."""
    torch.manual_seed(21)
    k, P, HW, Cc, G = 2, 2, 256, 320, 32
    nb = k * P
    nfl = N.raw().pbk_gn_tmp_floats
    nfl.restype = C.c_size_t
    tmp = torch.empty(nfl(HW, Cc, G, nb) + 64, device="cuda")
    gamma, beta = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    # one buffer holds both problems' primal tensors at stride `ps` floats: [x | mean | rstd] per problem
    ps = HW * Cc + 2 * HW + 64
    prim = torch.randn(P, ps, device="cuda")
    prim[:, HW * Cc + HW:HW * Cc + 2 * HW] = prim[:, HW * Cc + HW:HW * Cc + 2 * HW].abs() + 0.5     # rstd > 0
    xp = lambda s: C.c_void_p(prim[s].data_ptr())
    mp = lambda s: C.c_void_p(prim[s].data_ptr() + 4 * HW * Cc)
    rp = lambda s: C.c_void_p(prim[s].data_ptr() + 4 * (HW * Cc + HW))
    t16 = torch.randn(nb, HW, Cc, device="cuda").half()
    for fn in ("gn", "ln"):
        for mode in (0, 1):
            both = torch.empty_like(t16); sep = torch.empty_like(t16)
            if fn == "gn":
                call = lambda s0, tt, n, oo, ks, st_: _ok(N.leaf("pbk_gn_lin")(xp(s0), mp(s0), rp(s0), _p(gamma), _p(beta), HW, Cc, G, 1, _p(tt), n, mode,
                                                                            _p(oo), C.c_float(0), IO16, _p(tmp), ks, C.c_long(st_), _st()))
            else:
                call = lambda s0, tt, n, oo, ks, st_: _ok(N.leaf("pbk_ln_lin")(xp(s0), mp(s0), rp(s0), _p(gamma), C.c_long(HW), Cc, _p(tt), n, mode, _p(oo),
                                                                            C.c_float(0), IO16, ks, C.c_long(st_), _st()))
            call(0, t16, nb, both, k, ps)
            for s0 in range(P):
                call(s0, t16[s0 * k:(s0 + 1) * k], k, sep[s0 * k:(s0 + 1) * k], 0, 0)
            assert torch.equal(both, sep), (fn, mode)
    # GEGLU
    Fd = 256
    psg = HW * 2 * Fd + 32
    hp = torch.randn(P, psg, device="cuda")
    dh = torch.randn(nb, HW, 2 * Fd, device="cuda").half()
    both = torch.empty(nb, HW, Fd, device="cuda", dtype=torch.float16); sep = torch.empty_like(both)
    _ok(N.leaf("pbk_geglu_jvp")(_p(hp), C.c_long(HW), _p(dh), nb, Fd, _p(both), IO16, k, C.c_long(psg), _st()))
    for s0 in range(P):
        _ok(N.leaf("pbk_geglu_jvp")(_p(hp[s0]), C.c_long(HW), _p(dh[s0 * k:(s0 + 1) * k]), k, Fd, _p(sep[s0 * k:(s0 + 1) * k]), IO16, 0, C.c_long(0), _st()))
    assert torch.equal(both, sep)
    gy = torch.randn(nb, HW, Fd, device="cuda").half()
    both = torch.empty(nb, HW, 2 * Fd, device="cuda", dtype=torch.float16); sep = torch.empty_like(both)
    _ok(N.leaf("pbk_geglu_vjp")(_p(hp), C.c_long(HW), _p(gy), nb, Fd, _p(both), IO16, k, C.c_long(psg), _st()))
    for s0 in range(P):
        _ok(N.leaf("pbk_geglu_vjp")(_p(hp[s0]), C.c_long(HW), _p(gy[s0 * k:(s0 + 1) * k]), k, Fd, _p(sep[s0 * k:(s0 + 1) * k]), IO16, 0, C.c_long(0), _st()))
    assert torch.equal(both, sep)
    # fused attention, JVP role: primal operands (X16 = [N][3C] halves, P16, Vt16, O) per problem at a byte stride
    Mr = Nc = 256; d = 40; nh = 2; Ch = nh * d
    def blob():
        return dict(X=torch.randn(Mr, 3 * Ch, device="cuda").half(), P=(torch.softmax(torch.randn(nh, Mr, Nc, device="cuda") * 2, -1) * 16).half(),
                    Vt=torch.randn(nh, d, Nc, device="cuda").half(), O=torch.randn(Mr, Ch, device="cuda"))
    blobs = [blob() for _ in range(P)]
    sizes = {kk: v.numel() * v.element_size() for kk, v in blobs[0].items()}
    offs, top = {}, 0
    for kk in ("X", "P", "Vt", "O"):
        offs[kk] = top; top += (sizes[kk] + 255) // 256 * 256
    cache = torch.zeros(P, top, device="cuda", dtype=torch.uint8)
    for s0 in range(P):
        for kk in offs:
            cache[s0, offs[kk]:offs[kk] + sizes[kk]] = blobs[s0][kk].view(torch.uint8).reshape(-1)
    dqkv = torch.randn(nb, Mr, 3 * Ch, device="cuda").half()
    dVt = torch.randn(nb, nh, d, Nc, device="cuda").half()
    def run(s0, n, tq, tv, out, ks, stride):
        base = cache[s0].data_ptr()
        a = N.PbAttnLin()
        a.Mr, a.Nc, a.d, a.nb, a.nh, a.nseg = Mr, Nc, d, n, nh, 2
        q0 = a.seg[0]; q1 = a.seg[1]
        q0.A, q0.lda, q0.sAb, q0.sAh, q0.B, q0.ldb, q0.sBb, q0.sBh = tq.data_ptr(), 3 * Ch, Mr * 3 * Ch, d, base + offs["X"] + 2 * Ch, 3 * Ch, 0, d
        q1.A, q1.lda, q1.sAb, q1.sAh, q1.B, q1.ldb, q1.sBb, q1.sBh = base + offs["X"], 3 * Ch, 0, d, tq.data_ptr() + 2 * Ch, 3 * Ch, Mr * 3 * Ch, d
        a.alpha1, a.alpha2 = d ** -0.5, 1.0
        a.Pm, a.ldp, a.sPh = base + offs["P"], Nc, Mr * Nc
        a.want_rsum, a.O, a.ldo = 1, base + offs["O"], Ch
        a.C1, a.ldc, a.sCh = base + offs["Vt"], Nc, d * Nc
        a.C2, a.ldc2, a.sC2h, a.sC2b = tv.data_ptr(), Nc, d * Nc, nh * d * Nc
        a.D, a.ldd, a.sDb = out.data_ptr(), Ch, Mr * Ch
        a.p16, a.p_scale, a.s16, a.k_slot, a.p_stride = 1, 16.0, 1, ks, stride
        _ok(N.leaf("pbk_attn_lin")(C.byref(a), _st()))
    both = torch.empty(nb, Mr, Ch, device="cuda", dtype=torch.float16); sep = torch.empty_like(both)
    run(0, nb, dqkv, dVt, both, k, top)
    for s0 in range(P):
        run(s0, k, dqkv[s0 * k:(s0 + 1) * k], dVt[s0 * k:(s0 + 1) * k], sep[s0 * k:(s0 + 1) * k], 0, 0)
    assert torch.equal(both, sep)
    # attn_delta
    go = torch.randn(nb, Mr, Ch, device="cuda").half()
    both = torch.empty(nb, nh, Mr, device="cuda"); sep = torch.empty_like(both)
    obase = cache[0].data_ptr() + offs["O"]
    _ok(N.leaf("pbk_attn_delta")(_p(go), C.c_long(Ch), C.c_void_p(obase), C.c_long(Ch), nb, Mr, nh, d, _p(both), 4, k, C.c_long(top // 4), _st()))
    for s0 in range(P):
        _ok(N.leaf("pbk_attn_delta")(_p(go[s0 * k:(s0 + 1) * k]), C.c_long(Ch), C.c_void_p(cache[s0].data_ptr() + offs["O"]), C.c_long(Ch), k, Mr, nh, d,
                                     _p(sep[s0 * k:(s0 + 1) * k]), 4, 0, C.c_long(0), _st()))
    assert torch.equal(both, sep)


@pytest.mark.parametrize("Mr,Nc,d,nh,nb,case", [
    (1024, 1024, 16, 4, 3, "jvp_qkv"), (1024, 1024, 40, 2, 5, "jvp_qkv"), (512, 512, 64, 2, 5, "jvp_qkv"),
    (1024, 1024, 16, 4, 3, "vjp_b"), (1024, 1024, 40, 2, 5, "vjp_b"), (512, 576, 64, 2, 4, "vjp_b"),
    (1024, 1024, 40, 2, 5, "vjp_a"), (1024, 77, 40, 2, 5, "cross"), (1024, 1024, 32, 4, 7, "jvp"),
    (2048, 2048, 40, 8, 5, "vjp_b"), (2048, 2048, 40, 8, 5, "jvp_qkv")])
def test_attn_lin_engine_roles(Mr, Nc, d, nh, nb, case):
    """The fused attention linearisation in the operand layouts of the engine (tests/attn_cases.py), through the column-batched
    kernel (pb_attn16_sm100.cu): column groups of 5 / 3 + 2 / 4 + 3 tangents, the VJP-B role with a primal A and a per-tangent B
    operand, and 2048-token launches of 256 CTAs (more than one wave: S / T / per-column rings walked by two warps each)."""
    from tests.attn_cases import build, reference
    torch.manual_seed(3)
    a, t, _, scale = build(Mr, Nc, d, nb, nh, case)
    _ok(N.leaf("pbk_attn_lin")(C.byref(a), _st()))
    ref, e2 = reference(t, Mr, Nc, d, nb, nh, case, scale)
    assert rel(t["D"].float(), ref) < 8e-4, rel(t["D"].float(), ref)
    if e2 is not None:
        assert rel(t["D2"].float(), e2) < 6e-4, rel(t["D2"].float(), e2)
