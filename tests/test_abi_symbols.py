"""The C-ABI shared library loads on a CPU-only box and exports every entry point include/pullback_b200.h declares
(no compute calls here: the product path needs a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pullback_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_]*\s*\**\s+\**(pb_[a-z_0-9]+)\s*\(", src, flags=re.M)
    return sorted(set(names))


def test_header_declares_the_documented_entry_points():
    names = _declared()
    for must in ("pb_create", "pb_destroy", "pb_plan", "pb_bind_weights", "pb_set_point", "pb_jvp", "pb_vjp",
                 "pb_orthonormalize", "pb_pullback", "pb_pullback_host", "pb_last_error", "pb_backend"):
        assert must in names, must


def _declared_leaf():
    """The leaf-kernel interface (include/pb_kernels.h, include/pb_gemm.h): exported for kernel unit tests / micro-benchmarks."""
    names = []
    for f in ("pb_kernels.h", "pb_gemm.h"):
        src = open(os.path.join(ROOT, "include", f)).read()
        src = re.sub(r"//[^\n]*", "", src)
        names += re.findall(r"\b(pbk_[a-z_0-9]+|pb_gemm_tune[a-z_0-9]*)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    import subprocess
    from diffusion_pullback_b200.build import build
    path = build()
    lib = C.CDLL(path)
    missing = [n for n in _declared() + _declared_leaf() if not hasattr(lib, n)]
    assert not missing, missing
    # ... and nothing else: every exported pb* symbol is declared in include/*.h
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
    exported = sorted({l.split()[-1] for l in out.splitlines() if re.search(r" [TDB] (pb_|pbk_)", l)})
    undeclared = [n for n in exported if n not in set(_declared() + _declared_leaf())]
    assert not undeclared, undeclared
    lib.pb_backend.restype = C.c_char_p
    assert lib.pb_backend() == b"cuda-sm100a"


def test_product_loader_refuses_the_host_double():
    from diffusion_pullback_b200 import _native as N
    from tests.hostsim.build import build
    double = C.CDLL(build())
    double.pb_backend.restype = C.c_char_p
    assert double.pb_backend() != b"cuda-sm100a"
    assert N.lib().pb_backend() == b"cuda-sm100a"


def test_product_api_has_no_cpu_path():
    import torch
    import diffusion_pullback_b200 as PB
    from oracle import unet_torch as UT
    m = PB.patch_unet(UT.build_unet("uncond_tiny"))
    x, t, _ = UT.synthetic_inputs("uncond_tiny")
    with pytest.raises(RuntimeError):
        m.local_encoder_pullback_xt(x, t, op="mid", block_idx=0, pca_rank=2, min_iter=1, max_iter=1)


def test_plan_of_the_headline_workload_host_only():
    """pb_create / pb_plan / pb_plan_summary are host logic: on the product library, without a GPU, the SD-v1.5 mid-block plan
    must be the all-fp16 tangent plan (every GEMM of both passes reads and writes halves, nothing converted on the way) with
    every >= 512-token attention layer on the fused kernel with fp16 operands."""
    from diffusion_pullback_b200 import _native as N
    from diffusion_pullback_b200 import synthetic as SY
    from diffusion_pullback_b200.engine import unet_config
    L = N.lib()
    cfg = unet_config(SY.SyntheticUNet("sd15"))
    c = N.PbUnetCfg()
    c.kind, c.in_channels, c.n_levels = cfg["kind"], cfg["in_channels"], len(cfg["block_out_channels"])
    for i, v in enumerate(cfg["block_out_channels"]):
        c.block_out_channels[i], c.down_has_attn[i], c.up_has_attn[i], c.heads[i] = v, cfg["down_has_attn"][i], cfg["up_has_attn"][i], cfg["heads"][i]
    c.layers_per_block, c.cross_attention_dim = cfg["layers_per_block"], cfg["cross_attention_dim"]
    c.norm_num_groups, c.norm_eps = cfg["norm_num_groups"], cfg["norm_eps"]
    c.flip_sin_to_cos, c.freq_shift, c.downsample_padding = cfg["flip_sin_to_cos"], cfg["freq_shift"], cfg["downsample_padding"]
    h = C.c_void_p()
    assert L.pb_create(C.byref(c), C.byref(h)) == 0
    try:
        sizes = N.PbSizes()
        assert L.pb_plan(h, 64, 64, 0, 0, 5, 77, C.byref(sizes)) == 0
        assert (sizes.n_in, sizes.n_out) == (4 * 64 * 64, 1280 * 8 * 8)
        info = N.PbPlanInfo()
        assert L.pb_plan_summary(h, C.byref(info)) == 0
        # 10 resnets (2 per level on 4 levels + 2 in the mid block) x 2 convs = 20 stride-1 3x3 convs (the 3 downsamplers are im2col GEMMs)
        assert info.n_conv3x3 == 20
        assert info.n_gemm_f16_jvp == info.n_gemm_f16_vjp_stored == info.n_gemm_d16_jvp == info.n_gemm == 81
        assert info.n_gemm_f16_vjp_converted == 0                             # no fp32 -> fp16 staging copies left
        assert info.n_attn == 14                                              # 7 transformer blocks x (self + cross)
        assert info.n_attn_fused_self == 4 and info.n_attn_fused_cross == 4   # the 64x64 and 32x32 levels (>= 512 tokens)
        assert info.n_attn_p16 == 8                                           # all of them with fp16 probabilities and score operands
    finally:
        L.pb_destroy(h)


def test_ctypes_descriptor_mirrors_match_the_library():
    """PbGemm / PbAttnLin as mirrored in _native.py have the size the product library (and the host double) were compiled with:
    a field added on one side only would make the C side read past the Python struct."""
    from diffusion_pullback_b200 import _native as N
    from diffusion_pullback_b200.build import build
    from tests.hostsim.build import build as build_hostsim
    for path in (build(), build_hostsim()):
        N.check_struct_layout(C.CDLL(path))
    short = type("Short", (C.Structure,), {"_fields_": N.PbGemm._fields_[:-1]})
    saved = N.PbGemm
    try:
        N.PbGemm = short                                       # a mirror that lags the header is refused
        with pytest.raises(RuntimeError):
            N.check_struct_layout(C.CDLL(build()))
    finally:
        N.PbGemm = saved
