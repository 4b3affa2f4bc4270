"""ctypes loader for libpullback_b200.so.  Fails loudly when the CUDA library is missing: there is
no fallback path (the product is the sm_100a library)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libpullback_b200.so")

PB_MAX_LEVELS = 8


class PbUnetCfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("in_channels", C.c_int32), ("n_levels", C.c_int32),
                ("block_out_channels", C.c_int32 * PB_MAX_LEVELS),
                ("down_has_attn", C.c_int32 * PB_MAX_LEVELS),
                ("up_has_attn", C.c_int32 * PB_MAX_LEVELS),
                ("heads", C.c_int32 * PB_MAX_LEVELS),
                ("layers_per_block", C.c_int32), ("cross_attention_dim", C.c_int32),
                ("norm_num_groups", C.c_int32), ("norm_eps", C.c_float),
                ("flip_sin_to_cos", C.c_int32), ("freq_shift", C.c_float),
                ("downsample_padding", C.c_int32)]


class PbTensorDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 4)]


class PbSizes(C.Structure):
    _fields_ = [("packed_weight_bytes", C.c_size_t), ("primal_cache_bytes", C.c_size_t),
                ("workspace_bytes", C.c_size_t), ("n_in", C.c_int64), ("n_out", C.c_int64),
                ("out_channels", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32),
                ("in_channels", C.c_int32), ("in_h", C.c_int32), ("in_w", C.c_int32), ("n_x", C.c_int64)]


class PbIterInfo(C.Structure):
    _fields_ = [("iters_done", C.c_int32), ("converged", C.c_int32), ("last_dist", C.c_float)]


class PbPlanInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("n_ops", "n_gemm", "n_conv3x3", "n_gemm_f16_jvp", "n_gemm_f16_vjp_stored",
                                         "n_gemm_f16_vjp_converted", "n_gemm_d16_jvp", "n_attn", "n_attn_fused_self",
                                         "n_attn_fused_cross", "n_attn_p16")]


class PbGemmSeg(C.Structure):
    _fields_ = [("A", C.c_void_p), ("lda", C.c_long), ("sAb", C.c_long), ("sAh", C.c_long),
                ("B", C.c_void_p), ("ldb", C.c_long), ("sBb", C.c_long), ("sBh", C.c_long),
                ("K", C.c_int)]


class PbGemm(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("nseg", C.c_int), ("seg", PbGemmSeg * 2),
                ("D", C.c_void_p), ("ldd", C.c_long), ("sDb", C.c_long), ("sDh", C.c_long),
                ("R", C.c_void_p), ("ldr", C.c_long), ("sRb", C.c_long), ("sRh", C.c_long),
                ("bias", C.c_void_p), ("alpha", C.c_float), ("beta", C.c_float),
                ("nb", C.c_int), ("nh", C.c_int), ("conv", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("round_tf32", C.c_int), ("precise", C.c_int),
                ("ws", C.c_void_p), ("ws_floats", C.c_long), ("ab_dtype", C.c_int), ("d_dtype", C.c_int),
                ("k_slot", C.c_int), ("p_stride", C.c_long),
                ("gg", C.c_void_p), ("gg_F", C.c_int), ("gg_rows_p", C.c_long), ("gg_k_slot", C.c_int), ("gg_p_stride", C.c_long)]


class PbAttnLin(C.Structure):
    _fields_ = [("Mr", C.c_int), ("Nc", C.c_int), ("d", C.c_int), ("nb", C.c_int), ("nh", C.c_int), ("nseg", C.c_int),
                ("seg", PbGemmSeg * 2), ("alpha1", C.c_float), ("alpha2", C.c_float), ("beta", C.c_float),
                ("Pm", C.c_void_p), ("ldp", C.c_long), ("sPh", C.c_long), ("delta", C.c_void_p), ("delta_mode", C.c_int),
                ("want_rsum", C.c_int), ("O", C.c_void_p), ("ldo", C.c_long), ("C1", C.c_void_p), ("ldc", C.c_long),
                ("sCh", C.c_long), ("D", C.c_void_p), ("ldd", C.c_long), ("sDb", C.c_long), ("R", C.c_void_p),
                ("ldr", C.c_long), ("sRb", C.c_long), ("round_tf32", C.c_int),
                ("C2", C.c_void_p), ("ldc2", C.c_long), ("sC2h", C.c_long), ("sC2b", C.c_long),
                ("D2", C.c_void_p), ("ldd2", C.c_long), ("sD2b", C.c_long), ("p16", C.c_int), ("p_scale", C.c_float),
                ("s16", C.c_int), ("k_slot", C.c_int), ("p_stride", C.c_long)]


_lib = None
_raw = None


def raw() -> C.CDLL:
    global _raw
    if _raw is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m diffusion_pullback_b200.build` "
                "(there is no CPU or PyTorch fallback for the pullback hot path)")
        _raw = C.CDLL(LIB_PATH)
        check_struct_layout(_raw)
    return _raw


def check_struct_layout(L) -> None:
    """The ctypes mirrors of PbGemm / PbAttnLin must have the size the library was compiled with (a field added on one side only
    would make the C side read past the Python struct)."""
    gb, ab = C.c_int(0), C.c_int(0)
    L.pbk_struct_sizes(C.byref(gb), C.byref(ab))
    if (gb.value, ab.value) != (C.sizeof(PbGemm), C.sizeof(PbAttnLin)):
        raise RuntimeError(f"descriptor layout mismatch: library PbGemm / PbAttnLin = {gb.value} / {ab.value} bytes, "
                           f"ctypes mirrors = {C.sizeof(PbGemm)} / {C.sizeof(PbAttnLin)}; rebuild the library or update _native.py")


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = raw()
        _declare(_lib)
        if _lib.pb_backend().decode() != "cuda-sm100a":
            raise RuntimeError("libpullback_b200.so is not the sm_100a product library")
    return _lib


def _declare(L):
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.pb_backend.restype = C.c_char_p
    L.pb_backend.argtypes = []
    L.pb_last_error.restype = C.c_char_p
    L.pb_last_error.argtypes = [vp]
    L.pb_create.argtypes = [C.POINTER(PbUnetCfg), C.POINTER(vp)]
    L.pb_destroy.argtypes = [vp]
    L.pb_destroy.restype = None
    L.pb_plan.argtypes = [vp, i32, i32, i32, i32, i32, i32, C.POINTER(PbSizes)]
    L.pb_bind_weights.argtypes = [vp, C.POINTER(PbTensorDesc), i32, vp, vp]
    L.pb_set_point.argtypes = [vp, vp, f32, vp, vp, vp, vp, vp]
    L.pb_jvp.argtypes = [vp, vp, i32, vp, vp]
    L.pb_set_slots.argtypes = [vp, i32, C.c_size_t]
    L.pb_set_slots.restype = C.c_int
    L.pb_select_slot.argtypes = [vp, i32]
    L.pb_select_slot.restype = C.c_int
    L.pb_vjp.argtypes = [vp, vp, i32, vp, vp]
    L.pb_orthonormalize.argtypes = [vp, vp, vp, i32, f32, vp, vp, vp, vp]
    L.pb_pullback.argtypes = [vp, vp, i32, i32, i32, f32, vp, vp, vp, C.POINTER(PbIterInfo), vp]
    L.pb_pullback_host.argtypes = [vp, vp, f32, vp, vp, i32, i32, i32, f32, vp, vp, vp,
                                   C.POINTER(PbIterInfo), vp]
    L.pb_pullback_host_slots.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, f32, vp, vp, vp, C.POINTER(PbIterInfo), vp]
    L.pb_pullback_host_slots.restype = C.c_int
    L.pb_kernel_launches.argtypes = [vp]
    L.pb_kernel_launches.restype = C.c_int64
    L.pb_weight_count.argtypes = [vp]
    L.pb_weight_count.restype = C.c_int
    L.pb_weight_info.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(i32), C.POINTER(i64)]
    L.pb_weight_info.restype = C.c_int
    L.pb_set_option.argtypes = [vp, C.c_char_p, C.c_int]
    L.pb_set_option.restype = C.c_int
    L.pb_plan_summary.argtypes = [vp, C.POINTER(PbPlanInfo)]
    L.pb_plan_summary.restype = C.c_int
    L.pb_decode_from.argtypes = [vp, vp, vp, vp]
    L.pb_decode_from.restype = C.c_int
    L.pb_ddim_step.argtypes = [vp, vp, f32, f32, vp, vp, i64, vp]
    L.pb_ddim_step.restype = C.c_int
    L.pb_lincomb3.argtypes = [vp, f32, vp, f32, vp, f32, vp, i64, vp]
    L.pb_lincomb3.restype = C.c_int
    L.pb_profile_begin.argtypes = [vp]
    L.pb_profile_begin.restype = C.c_int
    L.pb_profile_read.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)]
    L.pb_profile_read.restype = C.c_int
    for name in ("pb_create", "pb_plan", "pb_bind_weights", "pb_set_point", "pb_jvp", "pb_vjp",
                 "pb_orthonormalize", "pb_pullback", "pb_pullback_host"):
        getattr(L, name).restype = C.c_int


def leaf(name):
    """A pbk_* leaf kernel entry (tests only): returns an error string or None."""
    f = getattr(raw(), name)
    f.restype = C.c_char_p
    return f
