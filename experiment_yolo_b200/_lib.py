"""ctypes binding of libldconv_b200.so (C ABI: include/ldconv_b200.h).

The library is built in-tree by `make -C experiment_yolo_b200/csrc` (see __graft_entry__.build()).  There is NO
fallback: if the shared object is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libldconv_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_SILU, ACT_LEAKY01 = 0, 1, 2
IMPL_FFMA, IMPL_TCGEN05 = 1, 2
FLAG_FORCE_FFMA = 1
FLAG_GATHER_DIRECT = 2

_vp = ctypes.c_void_p
_i = ctypes.c_int
_ll = ctypes.c_longlong
_f = ctypes.c_float

# name -> (restype, argtypes); mirrors include/ldconv_b200.h one to one
SIGNATURES = {
    "ldconv_version": (_i, []),
    "ldconv_last_error": (ctypes.c_char_p, []),
    "ldconv_device_check": (_i, []),
    "ldconv_last_impl": (_i, []),
    "ldconv_set_flag": (_i, [_i, _i]),
    "ldconv_set_gather_miss_counter": (_i, [_vp]),
    "ldconv_p_n": (_i, [_i, _vp]),
    "ldconv_offset_conv_fwd": (_i, [_vp, _vp, _vp, _vp] + [_i] * 7 + [_vp]),
    "ldconv_offset_conv_tc_supported": (_i, [_i] * 4),
    "ldconv_offset_conv_tc_fwd": (_i, [_vp, _vp, _vp, _vp] + [_i] * 7 + [_vp]),
    "ldconv_offset_conv_s2d_supported": (_i, [_i] * 5),
    "ldconv_offset_conv_s2d_fwd": (_i, [_vp, _vp, _vp, _vp] + [_i] * 6 + [_vp]),
    "ldconv_gather_fwd": (_i, [_vp] * 6 + [_i] * 7 + [_vp]),
    "ldconv_gemm_fwd": (_i, [_vp] * 8 + [_i] * 5 + [_vp]),
    "ldconv_col_stats": (_i, [_vp, _vp, _vp, _ll, _i, _i, _vp]),
    "ldconv_bn_finalize": (_i, [_vp, _vp, _ll, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "ldconv_bn_act_apply": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _vp]),
    "ldconv_bn_act_bwd_reduce": (_i, [_vp] * 7 + [_ll, _i, _i, _i, _vp]),
    "ldconv_bn_act_bwd_apply": (_i, [_vp] * 8 + [_ll, _i, _i, _i, _i, _vp]),
    "ldconv_gemm_bwd_weight": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ldconv_gather_bwd": (_i, [_vp] * 6 + [_i] * 7 + [_vp]),
    "ldconv_bwd_acc16_supported": (_i, [_i] * 6),
    "ldconv_gather_bwd_acc16": (_i, [_vp] * 6 + [_i] * 6 + [_vp]),
    "ldconv_offset_conv_bwd": (_i, [_vp] * 6 + [_i] * 7 + [_vp]),
    "ldconv_offset_conv_bwd_workspace_bytes": (ctypes.c_size_t, [_i] * 7),
    "ldconv_offset_conv_bwd_tc": (_i, [_vp] * 7 + [ctypes.c_size_t] + [_i] * 7 + [_vp]),
    "ldconv_offset_conv_bwd_tc_acc16": (_i, [_vp] * 7 + [ctypes.c_size_t] + [_i] * 6 + [_vp]),
    "ldconv_conv1x1_bn_act_fwd": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _ll, _i, _i, _i, _i, _vp]),
    "ldconv_conv1x1_bn_act_fwd2": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _ll, _i, _i, _i, _i, _vp]),
    "ldconv_conv3x3_supported": (_i, [_i] * 4),
    "ldconv_conv3x3_bn_act_fwd": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _i] + [_i] * 8 + [_vp]),
    "ldconv_conv1x1_detect_fwd": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _i, _i, _i, _vp]),
    "ldconv_detect_decode": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _i, _i, _vp]),
    "ldconv_nms_workspace_bytes": (ctypes.c_size_t, [_i, _i]),
    "ldconv_nms": (_i, [_vp, _vp, _vp, _vp, ctypes.c_size_t, _i, _i, _i, _f, _f, _i, _i, _i, _f, _i, _vp]),
    "ldconv_head_decode_rows": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _vp]),
    "ldconv_tal_metric": (_i, [_vp] * 8 + [_i] * 4 + [_f] * 3 + [_vp]),
    "ldconv_tal_assign": (_i, [_vp] * 12 + [_i] * 4 + [_f, _vp]),
    "ldconv_conv1x1_bn_act_maxup_fwd": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _i, _vp, _i, _vp, _i] + [_i] * 7 + [_vp]),
    "ldconv_conv1x1_bn_act_packed_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _ll, _i, _i, _i, _i, _i, _vp]),
    "ldconv_upsample_nearest": (_i, [_vp, _i, _vp, _i] + [_i] * 6 + [_vp]),
    "ldconv_ssff_max_fwd": (_i, [_vp] * 4 + [ctypes.c_longlong, _i, _i, _vp]),
    "ldconv_ssff_max_bwd": (_i, [_vp] * 5 + [ctypes.c_longlong, _i, _i, _vp]),
    "ldconv_upsample_nearest_bwd": (_i, [_vp, _i, _vp, _i] + [_i] * 6 + [_vp]),
    "ldconv_add_nhwc": (_i, [_vp, _vp, _i, _vp, _i, _ll, _i, _i, _vp]),
    "ldconv_scalseq_tail": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i] + [_i] * 9 + [_vp]),
    "ldconv_image_u8_to_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    "ldconv_sppf_pools": (_i, [_vp, _vp, _vp, _vp] + [_i] * 7 + [_vp]),
    "ldconv_gather_gemm_supported": (_i, [_i] * 9),
    "ldconv_gather_gemm_fwd": (_i, [_vp] * 7 + [_i] * 10 + [_vp]),
    "ldconv_debug_onepass_trace": (_i, [_vp]),
    "ldconv_debug_l0_variant": (_i, [_i]),
    "ldconv_debug_l0_trace": (_i, [_vp]),
    "ldconv_onepass_supported": (_i, [_i] * 9),
    "ldconv_onepass_fwd": (_i, [_vp] * 8 + [_i, _vp] + [_i] * 9 + [_vp]),
    "ldconv_fused_supported": (_i, [_i] * 8),
    "ldconv_fused_fwd": (_i, [_vp] * 9 + [_i] * 9 + [_vp]),
}

_lib = None

# successful C-ABI compute calls by entry point since the last reset (bench.py's `gpu_launches` bookkeeping): every entry
# point below launches at least one kernel of this library per call
call_counts: dict = {}


def load() -> ctypes.CDLL:
    """Load the CUDA library.  Raises if it has not been built: the product path has no other implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C experiment_yolo_b200/csrc` "
                "(or __graft_entry__.build()).  experiment_yolo_b200 has no CPU / eager fallback.")
        import torch  # noqa: F401  (loads libcudart.so.12 into the process before our library needs it)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ldconv_last_error()
        raise RuntimeError(f"libldconv_b200 {what} failed (code {rc}): {msg.decode() if msg else ''}")
    call_counts[what] = call_counts.get(what, 0) + 1


def p_n_table(N: int):
    """conv.py:413-432 through the library (host-side helper) -> list of 2N ints, rows then columns."""
    buf = (ctypes.c_int32 * (2 * N))()
    check(load().ldconv_p_n(N, ctypes.cast(buf, _vp)), "ldconv_p_n")
    return list(buf)
