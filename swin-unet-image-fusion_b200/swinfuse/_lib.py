"""ctypes binding of libswinfuse.so (include/swinfuse.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``.  There is no
fallback of any kind: if the shared object is missing or a call fails, a Python exception is
raised.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_DIR, "libswinfuse.so")

SF_PREC_FP32 = 0
SF_PREC_BF16 = 1

_f = C.c_void_p  # every tensor pointer travels as an untyped device address


class SwinFuseError(RuntimeError):
    pass


class WindowAttnParams(C.Structure):
    _fields_ = [(n, _f) for n in (
        "q_src", "kv_src", "residual", "out", "ln_q_gamma", "ln_q_beta", "ln_kv_gamma", "ln_kv_beta",
        "wq", "bq", "wk", "bk", "wv", "bv", "wo", "bo", "bias_table")] + [
        (n, C.c_int) for n in ("B", "Hp", "Wp", "C", "num_heads", "head_dim", "wsh", "wsw", "shift")] + [
        ("ln_eps", C.c_float), ("precision", C.c_int), ("packed", _f)]


class WindowAttnBwdParams(C.Structure):
    _fields_ = [("fwd", WindowAttnParams)] + [(n, _f) for n in (
        "gout", "g_q_src", "g_kv_src", "g_ln_q_gamma", "g_ln_q_beta", "g_ln_kv_gamma", "g_ln_kv_beta",
        "g_wq", "g_bq", "g_wk", "g_bk", "g_wv", "g_bv", "g_wo", "g_bo", "g_bias_table", "add_to_g_q_src")]


class MlpParams(C.Structure):
    _fields_ = [(n, _f) for n in ("in_", "residual", "out", "ln_gamma", "ln_beta", "w1", "b1", "w2", "b2")] + [
        ("M", C.c_longlong), ("C", C.c_int), ("hidden", C.c_int), ("ln_eps", C.c_float), ("precision", C.c_int),
        ("packed", _f)]


class MlpBwdParams(C.Structure):
    _fields_ = [("fwd", MlpParams)] + [(n, _f) for n in (
        "gout", "g_in", "g_ln_gamma", "g_ln_beta", "g_w1", "g_b1", "g_w2", "g_b2", "add_to_g_in")]


class PatchParams(C.Structure):
    _fields_ = [(n, _f) for n in ("in_", "out", "w", "b", "ln_gamma", "ln_beta")] + [
        (n, C.c_int) for n in ("B", "H", "W", "Cin", "Cout", "mh", "mw", "encoder")] + [
        ("ln_eps", C.c_float), ("precision", C.c_int), ("packed", _f)]


class PatchBwdParams(C.Structure):
    _fields_ = [("fwd", PatchParams)] + [(n, _f) for n in ("gout", "g_in", "g_w", "g_b", "g_ln_gamma", "g_ln_beta")]


class HeadParams(C.Structure):
    _fields_ = [(n, _f) for n in ("x", "y", "out", "w1", "b1", "bn_gamma", "bn_beta", "running_mean", "running_var",
                                  "save_mean", "save_invstd", "w2", "b2")] + [
        (n, C.c_int) for n in ("B", "H", "W", "ksize", "training")] + [
        ("bn_eps", C.c_float), ("bn_momentum", C.c_float)]


class HeadBwdParams(C.Structure):
    _fields_ = [("fwd", HeadParams)] + [(n, _f) for n in (
        "gout", "g_x", "g_y", "g_w1", "g_b1", "g_bn_gamma", "g_bn_beta", "g_w2", "g_b2")]


class FusionLossParams(C.Structure):
    _fields_ = [(n, _f) for n in ("fusion", "ir", "vis", "loss", "total", "g_fusion")] + [
        (n, C.c_int) for n in ("B", "H", "W", "clamp01")] + [
        (n, C.c_float) for n in ("w_ir", "ssim_scale", "texture_scale", "intensity_scale", "r_ssim", "r_texture",
                                 "r_intensity")]


class ProfileEntry(C.Structure):
    _fields_ = [("name", C.c_char * 48), ("launches", C.c_longlong), ("total_ms", C.c_double), ("flops", C.c_double),
                ("bytes", C.c_double)]


_i, _ll, _fl, _sz, _d = C.c_int, C.c_longlong, C.c_float, C.c_size_t, C.c_double

# name -> (restype, argtypes): every symbol include/swinfuse.h declares
SIGNATURES = {
    "sf_abi_version": (_i, []),
    "sf_last_error": (C.c_char_p, []),
    "sf_launch_count": (_ll, []),
    "sf_reset_launch_count": (None, []),
    "sf_profile_enable": (_i, [_i]),
    "sf_profile_summary": (_i, [C.POINTER(ProfileEntry), _i]),
    "sf_nchw_to_nhwc": (_i, [_f, _f, _i, _i, _i, _i, _f]),
    "sf_nhwc_to_nchw": (_i, [_f, _f, _i, _i, _i, _i, _f]),
    "sf_bgr_to_ycrcb": (_i, [_f, _f, _f, _i, _i, _i, _f]),
    "sf_ycrcb_to_rgb": (_i, [_f, _f, _f, _i, _i, _i, _f]),
    "sf_pad_reflect": (_i, [_f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "sf_pad_reflect_bwd": (_i, [_f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "sf_crop": (_i, [_f, _f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "sf_crop_bwd": (_i, [_f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "sf_patch_merge": (_i, [_f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "sf_patch_unmerge": (_i, [_f, _f, _i, _i, _i, _i, _i, _i, _f]),
    "sf_window_partition": (_i, [_f, _f, _i, _i, _i, _i, _i, _i, _i, _f]),
    "sf_window_reverse": (_i, [_f, _f, _i, _i, _i, _i, _i, _i, _i, _f]),
    "sf_shift_mask": (_i, [_f, _i, _i, _i, _i, _f]),
    "sf_relative_position_bias": (_i, [_f, _f, _i, _i, _f]),
    "sf_layernorm": (_i, [_f, _f, _f, _f, _ll, _i, _fl, _i, _f]),
    "sf_window_attn_workspace_bytes": (_sz, [C.POINTER(WindowAttnParams)]),
    "sf_window_attn_fwd": (_i, [C.POINTER(WindowAttnParams), _f, _sz, _f]),
    "sf_window_attn_packed_bytes": (_sz, [C.POINTER(WindowAttnParams)]),
    "sf_window_attn_pack": (_i, [C.POINTER(WindowAttnParams), _f, _sz, _f]),
    "sf_mlp_packed_bytes": (_sz, [C.POINTER(MlpParams)]),
    "sf_mlp_pack": (_i, [C.POINTER(MlpParams), _f, _sz, _f]),
    "sf_patch_packed_bytes": (_sz, [C.POINTER(PatchParams)]),
    "sf_patch_pack": (_i, [C.POINTER(PatchParams), _f, _sz, _f]),
    "sf_window_attn_bwd_workspace_bytes": (_sz, [C.POINTER(WindowAttnBwdParams)]),
    "sf_window_attn_bwd": (_i, [C.POINTER(WindowAttnBwdParams), _f, _sz, _f]),
    "sf_mlp_workspace_bytes": (_sz, [C.POINTER(MlpParams)]),
    "sf_mlp_fwd": (_i, [C.POINTER(MlpParams), _f, _sz, _f]),
    "sf_mlp_bwd_workspace_bytes": (_sz, [C.POINTER(MlpBwdParams)]),
    "sf_mlp_bwd": (_i, [C.POINTER(MlpBwdParams), _f, _sz, _f]),
    "sf_patch_workspace_bytes": (_sz, [C.POINTER(PatchParams)]),
    "sf_patch_fwd": (_i, [C.POINTER(PatchParams), _f, _sz, _f]),
    "sf_patch_bwd_workspace_bytes": (_sz, [C.POINTER(PatchBwdParams)]),
    "sf_patch_bwd": (_i, [C.POINTER(PatchBwdParams), _f, _sz, _f]),
    "sf_head_workspace_bytes": (_sz, [C.POINTER(HeadParams)]),
    "sf_head_fwd": (_i, [C.POINTER(HeadParams), _f, _sz, _f]),
    "sf_head_bwd_workspace_bytes": (_sz, [C.POINTER(HeadBwdParams)]),
    "sf_head_bwd": (_i, [C.POINTER(HeadBwdParams), _f, _sz, _f]),
    "sf_add": (_i, [_f, _f, _f, _ll, _f]),
    "sf_fusion_loss_workspace_bytes": (_sz, [C.POINTER(FusionLossParams)]),
    "sf_fusion_loss": (_i, [C.POINTER(FusionLossParams), _f, _sz, _f]),
    "sf_scale_by_scalar": (_i, [_f, _f, _f, _ll, _f]),
    "sf_adam_step": (_i, [_f, _f, _f, _f, _ll, _d, _d, _d, _d, _i, _d, _f]),
}

_lib = None


def load() -> C.CDLL:
    """Load libswinfuse.so (once).  Raises SwinFuseError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SwinFuseError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            f"(or `make -C {os.path.join(_PKG_DIR, 'csrc')}`).  There is no CPU / PyTorch fallback.")
    import torch  # noqa: F401  (makes sure libcudart from the torch wheel is already mapped)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.restype = res
        fn.argtypes = args
    if lib.sf_abi_version() != 1:
        raise SwinFuseError(f"libswinfuse ABI version {lib.sf_abi_version()} != 1")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().sf_last_error().decode(errors="replace")
        raise SwinFuseError(f"{what} failed (status {rc}): {msg}")
