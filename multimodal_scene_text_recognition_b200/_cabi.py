"""ctypes binding of include/mstr_b200.h (the drop-in boundary).

There is deliberately no fallback: if ``libmstr_b200.so`` is missing or a call fails, an exception
is raised.  Nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_int64, c_void_p, c_char_p

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "libmstr_b200.so")

MSTR_F32, MSTR_BF16 = 0, 1

_P = c_void_p
_I = c_int

# name -> (restype, argtypes); mirrors include/mstr_b200.h one to one
SIGNATURES = {
    "mstr_version": (c_int, []),
    "mstr_last_error_string": (c_char_p, []),
    "mstr_launch_count": (c_int64, []),
    "mstr_copy_h2d_async": (c_int, [_P, _P, c_int64, _P]),
    "mstr_copy_d2h_async": (c_int, [_P, _P, c_int64, _P]),
    "mstr_tps_rectify": (c_int, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "mstr_conv3x3_first": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "mstr_conv_gemm_f32": (c_int, [_P, _P, _P, _P, _P, _P] + [_I] * 12 + [_P]),
    "mstr_conv_gemm_bf16": (c_int, [_P, _P, _P, _P, _P, _P, _P] + [_I] * 12 + [_P]),
    "mstr_conv_gemm_bf16_dual": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P] + [_I] * 13 + [_P]),
    "mstr_debug_gemm_route": (c_int, [_I]),
    "mstr_conv_gemm_tf32": (c_int, [_P, _P, _P, _P, _P, _P, _P] + [_I] * 12 + [_P]),
    "mstr_maxpool_nhwc": (c_int, [_P, _P] + [_I] * 11 + [_P]),
    "mstr_maxpool2x2_ring_bf16": (c_int, [_P, _P, _I, _I, _I, _I, _P]),
    "mstr_conv3x3_flat_bf16": (c_int, [_P, _P, _P, _P, _P] + [_I] * 6 + [_P]),
    "mstr_conv3x3_flat_pool_bf16": (c_int, [_P, _P, _P, _P, _P] + [_I] * 6 + [_P]),
    "mstr_debug_flat_desc_mode": (c_int, [_I]),
    "mstr_mean_hw": (c_int, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "mstr_locnet_head": (c_int, [_P, _I, _P, _P, _P, _P, _P] + [_I] * 5 + [_P]),
    "mstr_mean_h": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "mstr_cast": (c_int, [_P, _P, _I, _I, c_int64, _P]),
    "mstr_embed_semantics": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "mstr_bilstm_recurrent": (c_int, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "mstr_bilstm_recurrent_bf16": (c_int, [_P, _I, _P, _P, _I, _I, _I, _P]),
    "mstr_set_bilstm_kernel": (c_int, [_I]),
    "mstr_attn_decode": (c_int, [_P] * 20 + [_I] * 6 + [_P]),
    "mstr_attn_decode_bf16_workspace_bytes": (c_int64, [_I, _I]),
    "mstr_attn_decode_bf16": (c_int, [_P] * 15 + [_I] * 6 + [_P]),
    "mstr_set_decode_persistent": (c_int, [_I]),
    "mstr_debug_decode_trace": (c_int, [_P, _I, _P, _I]),
    "mstr_resize_to_tensor_u8": (c_int, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "mstr_layer_norm": (c_int, [_P, _P, _P, _P, c_int64, _I, ctypes.c_float, _I, _P]),
    "mstr_add_pos_encoding": (c_int, [_P, _I, _P, _P, _I, _I, _I, _I, _P]),
    "mstr_embed_tokens": (c_int, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "mstr_mha": (c_int, [_P, _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "mstr_mha_step": (c_int, [_P, _I, _P, _P, _I, c_int64, _P, _P, _I, _P, _I, _I, _I, _I, _I, _P]),
    "mstr_pair_add_relu": (c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "mstr_relevance_reduce": (c_int, [_P, _P, ctypes.c_float, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "mstr_gelu": (c_int, [_P, _P, c_int64, _I, _P]),
    "mstr_argmax_tokens": (c_int, [_P, _I, _P, c_int64, _P, _I, _P, _I, _I, _I, _P]),
}

_lib = None


class MstrError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load the CUDA extension (once).  Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: the sm_100a CUDA extension is not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or "
                "`python -m multimodal_scene_text_recognition_b200.build`). There is no CPU fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so does not export the symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().mstr_last_error_string().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise MstrError("%s failed with status %d: %s" % (what or "mstr call", rc, last_error()))


def launch_count() -> int:
    return int(lib().mstr_launch_count())


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
