"""ctypes binding of ``libns2vc_b200.so`` (C-ABI declared in ``include/ns2vc_b200.h``).

There is no CPU fallback: if the library is missing every CUDA entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", os.environ.get("NS2VC_LIB_NAME", "libns2vc_b200.so"))   # (NS2VC_LIB_NAME: A/B builds during development)

MAX_LEVELS = 8


class UNetCfg(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int), ("latent_channels", C.c_int), ("out_channels", C.c_int), ("n_levels", C.c_int),
        ("block_out_channels", C.c_int * MAX_LEVELS), ("layers_per_block", C.c_int * MAX_LEVELS),
        ("down_has_attn", C.c_int * MAX_LEVELS), ("up_has_attn", C.c_int * MAX_LEVELS),
        ("num_heads", C.c_int), ("cross_attention_dim", C.c_int), ("norm_num_groups", C.c_int), ("norm_eps", C.c_float),
        ("time_scale_shift", C.c_int), ("add_embed_text", C.c_int), ("add_embed_heads", C.c_int),
        ("flip_sin_to_cos", C.c_int), ("freq_shift", C.c_float),
    ]


class PreCfg(C.Structure):
    _fields_ = [("phone_in", C.c_int), ("phone_hidden", C.c_int), ("phone_out", C.c_int), ("phone_layers", C.c_int),
                ("prompt_in", C.c_int), ("prompt_hidden", C.c_int), ("prompt_out", C.c_int), ("prompt_layers", C.c_int),
                ("ref_dim", C.c_int), ("ref_heads", C.c_int), ("n_heads", C.c_int), ("ffn_kernel", C.c_int)]


class DpmCoef(C.Structure):
    _fields_ = [("alpha_s", C.c_float), ("sigma_s", C.c_float), ("c_x", C.c_float), ("c_m", C.c_float),
                ("c_d", C.c_float), ("inv_r0", C.c_float), ("order", C.c_int)]


class UniPcCoef(C.Structure):
    _fields_ = [("alpha_t", C.c_float), ("sigma_t", C.c_float), ("c_x", C.c_float), ("c_m", C.c_float),
                ("ab", C.c_float), ("rk", C.c_float), ("rho0", C.c_float), ("rho1", C.c_float), ("corr_order", C.c_int),
                ("n_c_x", C.c_float), ("n_c_m", C.c_float), ("nab", C.c_float), ("nrk", C.c_float), ("pred_order", C.c_int)]


# symbol -> (restype, argtypes); also the export list checked by tests/test_abi.py
_P = C.c_void_p
SIGNATURES = {
    "ns2vc_last_error": (C.c_char_p, []),
    "ns2vc_build_info": (C.c_char_p, []),
    "ns2vc_unet_create": (C.c_int, [C.POINTER(UNetCfg), C.POINTER(_P)]),
    "ns2vc_unet_destroy": (None, [_P]),
    "ns2vc_unet_num_weights": (C.c_int, [_P]),
    "ns2vc_unet_weight_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "ns2vc_unet_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int, _P]),
    "ns2vc_unet_finalize": (C.c_int, [_P, _P]),
    "ns2vc_unet_workspace_bytes": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "ns2vc_unet_prepare_cond": (C.c_int, [_P, _P, C.c_longlong, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "ns2vc_unet_forward": (C.c_int, [_P, _P, C.c_longlong, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "ns2vc_unet_film_width": (C.c_int, [_P]),
    "ns2vc_unet_time_table_floats": (C.c_size_t, [_P, C.c_int]),
    "ns2vc_unet_time_table": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "ns2vc_unet_forward_film": (C.c_int, [_P, _P, C.c_longlong, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "ns2vc_dpm_step": (C.c_int, [_P, _P, _P, C.POINTER(DpmCoef), _P, _P, C.c_size_t, _P, _P]),
    "ns2vc_unipc_step": (C.c_int, [_P, _P, _P, _P, _P, C.POINTER(UniPcCoef), _P, _P, _P, C.c_size_t, _P, _P]),
    "ns2vc_mask_bias": (C.c_int, [_P, C.c_int, _P, _P]),
    "ns2vc_nearest_index": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "ns2vc_down_length": (C.c_int, [C.c_int]),
    "ns2vc_unet_num_taps": (C.c_int, [_P]),
    "ns2vc_unet_tap_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ns2vc_unet_set_tap": (C.c_int, [_P, C.c_int, _P]),
    "ns2vc_unet_plan_string": (C.c_char_p, [_P]),
    "ns2vc_unet_launch_count": (C.c_int, [_P]),
    "ns2vc_unet_set_profiling": (C.c_int, [_P, C.c_int]),
    "ns2vc_unet_set_trace": (C.c_int, [_P, _P, C.c_int]),
    "ns2vc_unet_set_attn_trace": (C.c_int, [_P, _P, C.c_int]),
    "ns2vc_unet_set_span_trace": (C.c_int, [_P, _P, C.c_int]),
    "ns2vc_unet_launch_kind": (C.c_int, [_P, C.c_int]),
    "ns2vc_profile_num_kinds": (C.c_int, []),
    "ns2vc_profile_kind_name": (C.c_char_p, [C.c_int]),
    "ns2vc_unet_profile_read": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "ns2vc_unet_profile_dump": (C.c_int, [_P, C.c_char_p]),
    "ns2vc_unet_profile_reset": (C.c_int, [_P]),
    # condition encoders (Pre_model)
    "ns2vc_pre_create": (C.c_int, [C.POINTER(PreCfg), C.POINTER(_P)]),
    "ns2vc_pre_destroy": (None, [_P]),
    "ns2vc_pre_num_weights": (C.c_int, [_P]),
    "ns2vc_pre_weight_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int)]),
    "ns2vc_pre_load_weight": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int, _P]),
    "ns2vc_pre_finalize": (C.c_int, [_P, _P]),
    "ns2vc_pre_workspace_bytes": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "ns2vc_pre_infer": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "ns2vc_pre_num_taps": (C.c_int, [_P]),
    "ns2vc_pre_tap_info": (C.c_int, [_P, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ns2vc_pre_set_tap": (C.c_int, [_P, C.c_int, _P]),
    "ns2vc_pre_launch_count": (C.c_int, [_P]),
}

_lib: Optional[C.CDLL] = None


class Ns2vcError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the shared library (once).  Raises loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Ns2vcError(
                f"{LIB_PATH} not found: the ns2vc_b200 CUDA extension is not built "
                "(run ./build.sh or __graft_entry__.build()); there is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().ns2vc_last_error()
        raise Ns2vcError((msg or b"unknown error").decode("utf-8", "replace") + f" (code {rc})")
