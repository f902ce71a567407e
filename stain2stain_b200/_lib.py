"""ctypes binding of the C-ABI library (include/s2s_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.  The library is loaded from
stain2stain_b200/lib/ (in-tree, so the driver sees which .so the process mapped).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import _build

_LIB = None


class S2SError(RuntimeError):
    pass


class ConvSrc(C.Structure):
    _fields_ = [("x", C.c_void_p), ("C", C.c_int), ("taps", C.c_int), ("stride", C.c_int)]


class PackJob(C.Structure):
    _fields_ = [("w", C.c_void_p), ("dst", C.c_void_p), ("Cout", C.c_int), ("Cin", C.c_int), ("taps", C.c_int),
                ("ci_begin", C.c_int), ("ci_count", C.c_int), ("ld_k", C.c_int), ("k_off", C.c_int),
                ("transpose_flip", C.c_int), ("fmt", C.c_int), ("mode", C.c_int)]


class GemmJob(C.Structure):
    _fields_ = [("A", C.c_void_p), ("B", C.c_void_p), ("C", C.c_void_p), ("bias", C.c_void_p), ("add", C.c_void_p),
                ("C2", C.c_void_p), ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("ldc", C.c_int), ("ld_add", C.c_int),
                ("sam", C.c_int), ("sak", C.c_int), ("sbk", C.c_int), ("sbn", C.c_int)]


class ConvNorm(C.Structure):
    _fields_ = [("coef", C.c_void_p), ("ld", C.c_int), ("off", C.c_int)]


_vp, _i, _f, _d, _u64, _ll = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_uint64, C.c_longlong

# name -> argtypes (restype is int unless listed in _RESTYPES).  Must mirror include/s2s_b200.h exactly.
SIGNATURES = {
    "s2s_last_error": [],
    "s2s_abi_version": [],
    "s2s_num_sms": [],
    "s2s_pack_conv_weight": [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp],
    "s2s_pack_conv_weight_mode": [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "s2s_upconv_supported": [_i, _i],
    "s2s_upconv_stat_tiles": [_i, _i, _i],
    "s2s_upconv_fwd": [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp],
    "s2s_upconv_dgrad": [_vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _i, _i, _vp],
    "s2s_upconv_wgrad": [_vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp],
    "s2s_upconv_unpack_wgrad": [_vp, _i, _i, _vp, _vp],
    "s2s_downconv_dgrad": [_vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _i, _i, _vp],
    "s2s_pack_tiles": [_i, _i, _i],
    "s2s_pack_conv_weight_multi": [_vp, _vp, _i, _vp],
    "s2s_conv_fwd": [C.POINTER(ConvSrc), _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _i, _vp],
    "s2s_conv_norm_fusable": [C.POINTER(ConvSrc), _i, _i],
    "s2s_conv_fwd_norm": [C.POINTER(ConvSrc), C.POINTER(ConvNorm), _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i,
                          _i, _i, _vp],
    "s2s_conv_stat_tiles": [_i, _i, _i],
    "s2s_conv_stat_tiles_for": [C.POINTER(ConvSrc), _i, _i, _i, _i],
    "s2s_conv_stat_tiles_geom": [C.POINTER(ConvSrc), _i, _i, _i, _i],
    "s2s_gn_coef_parts": [_vp, _i, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp, _vp],
    "s2s_conv_wgrad": [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _vp],
    "s2s_unpack_wgrad": [_vp, _i, _i, _i, _i, _i, _vp, _i, _i, _f, _vp],
    "s2s_patch27_pack": [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp],
    "s2s_gn_stats": [_vp, _i, _i, _i, _vp, _i, _i, _i, _vp],
    "s2s_gn_coef": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp],
    "s2s_gn_apply": [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _i, _i, _f, _u64, _vp, _i, _i, _vp],
    "s2s_gn_bwd_reduce": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _i, _f, _u64, _vp, _i, _i, _vp],
    "s2s_gn_bwd_reduce_x2": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _i, _f, _u64, _vp, _vp, _i, _i, _vp],
    "s2s_gn_bwd_coef": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "s2s_gn_chunks": [_i, _i],
    "s2s_gn_bwd_apply": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _f, _u64, _vp, _i, _i, _vp],
    "s2s_upsample2x": [_vp, _vp, _i, _i, _i, _i, _vp],
    "s2s_sumpool2x": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "s2s_zero_insert2x": [_vp, _vp, _i, _i, _i, _i, _vp],
    "s2s_channel_sum": [_vp, _ll, _i, _vp, _i, _vp],
    "s2s_fm_loss": [_vp, _vp, _vp, _ll, _vp, _vp, _vp],
    "s2s_patch_pack": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _i, _vp],
    "s2s_fm_loss_weighted": [_vp, _vp, _vp, _vp, _f, _i, _i, _i, _vp, _vp, _vp],
    "s2s_roi_charbonnier": [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp, _vp],
    "s2s_tile_prep": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "s2s_resample_u8": [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _vp],
    "s2s_denorm_u8": [_vp, _i, _i, _i, _vp, _vp],
    "s2s_convert16": [_vp, _vp, _ll, _i, _i, _vp],
    "s2s_nchw_f32_to_nhwc16": [_vp, _vp, _i, _i, _i, _i, _vp],
    "s2s_nhwc16_to_nchw_f32": [_vp, _vp, _i, _i, _i, _i, _vp],
    "s2s_bn_coef": [_vp, _i, _i, _i, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp],
    "s2s_bn_bwd_coef": [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "s2s_bn_fold_slices": [_i],
    "s2s_bn_fold": [_vp, _i, _i, _vp, _i, _vp],
    "s2s_bn_coef_sums": [_vp, _i, _i, _ll, _i, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp],
    "s2s_bn_bwd_coef_sums": [_vp, _i, _i, _ll, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "s2s_maxpool2x": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "s2s_maxpool2x_bwd": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "s2s_bilinear2x": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "s2s_bilinear2x_bwd": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "s2s_nchw_f32_to_nhwc16_pad": [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    "s2s_seg_loss_sums": [_vp, _vp, _i, _i, _i, _ll, _vp, _vp],
    "s2s_seg_loss_bwd": [_vp, _vp, _i, _i, _i, _ll, _vp, _f, _f, _f, _vp, _vp, _vp],
    "s2s_gn_apply_step": [_vp, _i, _i, _i, _vp, _i, _i, _vp, _vp, _i, _i, _f, _u64, _vp, _vp, _i, _i, _vp],
    "s2s_adam_multi_step": [_vp, _vp, _i, _d, _d, _d, _d, _d, _i, _vp, _d, _vp],
    "s2s_copy_multi": [_vp, _vp, _i, _vp],
    "s2s_head_conv": [_vp, _i, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _f, _i, _vp],
    "s2s_attn_supported": [_i],
    "s2s_attn_fwd": [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp],
    "s2s_attn_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "s2s_linear_max_jobs": [],
    "s2s_linear_multi": [C.POINTER(GemmJob), _i, _vp],
    "s2s_sum_parts_silu_bwd": [_vp, _i, _ll, _vp, _vp, _vp],
    "s2s_timestep_embedding": [_vp, _i, _i, _f, _vp, _vp],
    "s2s_adam_chunk": [],
    "s2s_adam_multi": [_vp, _vp, _i, _d, _d, _d, _d, _d, _i, _d, _vp],
}
_RESTYPES = {"s2s_last_error": C.c_char_p}


def lib_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if the .so is absent/stale and nvcc is present).  Raises if it cannot."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB_PATH
    override = os.environ.get("S2S_LIB_PATH")  # same-box A/B of two builds (scripts/gpu_ab_lib.sh); never set in production
    if override:
        if not os.path.exists(override):
            raise S2SError(f"S2S_LIB_PATH={override} does not exist")
        path, build_if_missing = override, False
    if build_if_missing and _build.needs_build():
        try:
            _build.build()
        except Exception as e:
            # a library older than its sources must never run silently: the kernels would not be the ones in the tree.
            # S2S_ALLOW_STALE_LIB=1 is the explicit override (e.g. a box without nvcc holding a deliberately prebuilt .so)
            if not os.path.exists(path):
                raise S2SError(f"libs2s_b200.so is missing and could not be built: {e}") from e
            if os.environ.get("S2S_ALLOW_STALE_LIB", "0") != "1":
                raise S2SError(f"{path} is older than csrc/ and the rebuild failed ({e}); fix the build or set "
                               f"S2S_ALLOW_STALE_LIB=1 to run the stale library knowingly") from e
    if not os.path.exists(path):
        raise S2SError(f"{path} not found: run `python -m stain2stain_b200._build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library drifted apart
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _LIB = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().s2s_last_error()
        raise S2SError(f"{what or 's2s call'} failed ({rc}): {msg.decode() if msg else '?'}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Refuses host tensors: the product path has no CPU fallback."""
    if t is None:
        return None
    if not t.is_cuda:
        raise S2SError("stain2stain_b200 kernels need CUDA tensors (no CPU fallback)")
    return t.data_ptr()
