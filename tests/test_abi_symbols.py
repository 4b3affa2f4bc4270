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


def test_library_exports_every_declared_symbol():
    from diffusion_pullback_b200.build import build
    lib = C.CDLL(build())
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
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
