"""ctypes binding of libmhada_b200.so (C ABI declared in include/mhada_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or the device is not a B200 the
calls raise.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmhada_b200.so")

F32, BF16, U8 = 0, 1, 2
ABI_VERSION = 13
VIT_MAX_LAYERS = 8
PROJ_Q, PROJ_KV = 1, 2
REUSE_FS_STATS = 1
LAYER_COSINE = 2
WOUT_BF16 = 4
ACT_SOFTMAX, ACT_COSINE = 0, 1


class AttnArgs(ctypes.Structure):
    """mhada_attn_args (include/mhada_b200.h)."""
    _fields_ = [
        ("dtype", c_int),
        ("B", c_int), ("H", c_int), ("Nc", c_int), ("Ns", c_int), ("dqk", c_int), ("dv", c_int),
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("x", c_void_p), ("out", c_void_p),
        ("ldq", c_int), ("ldk", c_int), ("ldv", c_int), ("ldx", c_int), ("ldo", c_int),
        ("x_mean", c_void_p), ("x_rstd", c_void_p), ("mu_v", c_void_p),
        ("q_mean", c_void_p), ("q_rstd", c_void_p), ("k_mean", c_void_p), ("k_rstd", c_void_p),
        ("kv_batch", c_int),
        ("activation", c_int),
        ("scratch", c_void_p), ("scratch_bytes", c_size_t),
    ]


class VitLayer(ctypes.Structure):
    """mhada_vit_layer (include/mhada_b200.h)."""
    _fields_ = [(n, c_void_p) for n in ("w_in", "w_out", "w_fc1", "w_fc2", "b_in", "b_out", "b_fc1", "b_fc2",
                                        "ln1_g", "ln1_b", "ln2_g", "ln2_b")]


class VitArgs(ctypes.Structure):
    """mhada_vit_args (include/mhada_b200.h)."""
    _fields_ = [
        ("img_dtype", c_int), ("img", c_void_p),
        ("B", c_int), ("Himg", c_int), ("Wimg", c_int), ("patch", c_int),
        ("D", c_int), ("F", c_int), ("heads", c_int), ("n_layers", c_int),
        ("w_patch", c_void_p), ("b_patch", c_void_p), ("pos", c_void_p),
        ("layers", VitLayer * 8),
        ("feat_f32", c_void_p * 8), ("feat_bf16", c_void_p * 8),
        ("ws", c_void_p), ("ws_bytes", c_size_t),
    ]


class ForlossArgs(ctypes.Structure):
    """mhada_forloss_args (include/mhada_b200.h)."""
    _fields_ = [("dtype", c_int), ("B", c_int), ("Nc", c_int), ("Ns", c_int), ("dqk", c_int), ("dv", c_int),
                ("c_x", c_void_p), ("s_x", c_void_p), ("c_1x", c_void_p), ("s_1x", c_void_p), ("out", c_void_p),
                ("ws", c_void_p), ("ws_bytes", c_size_t)]


class LayerBwdArgs(ctypes.Structure):
    """mhada_layer_bwd_args (include/mhada_b200.h)."""
    _fields_ = [("B", c_int), ("Nc", c_int), ("Ns", c_int), ("C", c_int), ("H", c_int),
                ("fc", c_void_p), ("fs", c_void_p), ("fcs", c_void_p),
                ("w_fgh", c_void_p), ("b_fgh", c_void_p), ("w_out", c_void_p), ("b_out", c_void_p),
                ("d_out", c_void_p),
                ("d_fc", c_void_p), ("d_fs", c_void_p), ("d_fcs", c_void_p),
                ("d_w_fgh", c_void_p), ("d_b_fgh", c_void_p), ("d_w_out", c_void_p), ("d_b_out", c_void_p),
                ("ws", c_void_p), ("ws_bytes", c_size_t)]


# name -> (restype, argtypes); kept in one table so tests can check the export list against the header
SIGNATURES = {
    "mhada_abi_version": (c_int, []),
    "mhada_profile_stage": (c_int, [c_int, POINTER(c_float), POINTER(c_int)]),
    "mhada_last_error": (c_char_p, []),
    "mhada_device_check": (c_int, []),
    "mhada_last_launch_count": (c_int, []),
    "mhada_total_launch_count": (ctypes.c_longlong, []),
    "mhada_profile_begin": (c_int, []),
    "mhada_profile_end": (c_int, [POINTER(c_float), POINTER(c_int)]),
    "mhada_in_stats_workspace": (c_size_t, [c_int, c_int, c_int]),
    "mhada_in_stats": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                               c_void_p]),
    "mhada_proj_workspace": (c_size_t, [c_int, c_int, c_int]),
    "mhada_proj": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                           c_size_t, c_void_p]),
    "mhada_attn": (c_int, [POINTER(AttnArgs), c_void_p]),
    "mhada_attn_cosine_scratch": (c_size_t, [c_int, c_int]),
    "mhada_debug_attn_trace": (c_int, [POINTER(AttnArgs), c_void_p, c_void_p]),
    "mhada_linear_workspace": (c_size_t, [c_int, c_int, c_int]),
    "mhada_linear": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                             c_void_p, c_size_t, c_void_p]),
    "mhada_forloss_workspace": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "mhada_forloss_forward": (c_int, [POINTER(ForlossArgs), c_void_p]),
    "mhada_layer_backward_workspace": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "mhada_layer_backward": (c_int, [POINTER(LayerBwdArgs), c_void_p]),
    "mhada_attn_bwd": (c_int, [c_int, c_int, c_int, c_int] + [c_void_p] * 15),
    "mhada_gemm_splitk_workspace": (c_size_t, [c_int, c_int, c_int]),
    "mhada_gemm_bf16_splitk": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                       c_size_t, c_void_p]),
    "mhada_transpose_bf16": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mhada_colsum_workspace": (c_size_t, [c_int, c_int]),
    "mhada_colsum": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "mhada_vit_workspace": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "mhada_vit_forward": (c_int, [POINTER(VitArgs), c_void_p]),
    "mhada_patch_im2col": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mhada_gemm_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "mhada_layernorm": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p]),
    "mhada_batch_attn": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mhada_batch_attn_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mhada_style_cache_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "mhada_style_precompute": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t,
                                       c_void_p, c_size_t, c_void_p]),
    "mhada_layer_forward_cached": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                           c_size_t, c_void_p]),
    "mhada_conv3x3_small": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p]),
    "mhada_conv3x3": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                              c_void_p]),
    "mhada_pad_reflect": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mhada_pad_reflect_bwd": (c_int, [c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mhada_layer_workspace": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "mhada_layer_forward": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None


class MhadaError(RuntimeError):
    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with code {code}: {msg}")
        self.code = code


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is not built (python -m mhada_style_transfer_b200.build). "
                "The MHAda hot path has no CPU / PyTorch fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        if l.mhada_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libmhada_b200.so ABI {l.mhada_abi_version()} != expected {ABI_VERSION}")
        _lib = l
    return _lib


def check(fn: str, code: int) -> None:
    if code != 0:
        raise MhadaError(fn, code, lib().mhada_last_error().decode("utf-8", "replace"))
