"""numpy front-end of the C oracle (oracle/ldconv_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs.
The product package (experiment_yolo_b200/) never imports this module and has no CPU fallback.

Every function restates /root/reference/ultralytics/nn/modules/conv.py:350-503 (class LDConv); the per-function
citations live in the C file.  Layouts are the reference's: NCHW fp32, offsets (B,2N,h,w) rows-then-columns.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libldconv_oracle.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_int = ctypes.c_int
_flt = ctypes.c_float


def build(force: bool = False) -> str:
    """Compile the C oracle with oracle/Makefile (gcc only) and return the .so path."""
    src = os.path.join(_HERE, "ldconv_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        L = _lib
        L.ldc_oracle_p_n.argtypes = [_int, _i64p]
        L.ldc_oracle_out_size.argtypes = [_int, _int]
        L.ldc_oracle_out_size.restype = _int
        L.ldc_oracle_offset_conv.argtypes = [_f32p, _f32p, _f32p, _f32p] + [_int] * 6
        L.ldc_oracle_grid.argtypes = [_f32p, _i64p, _i32p, _f32p, _f32p] + [_int] * 7
        L.ldc_oracle_sample.argtypes = [_f32p, _f32p, _i64p, _f32p] + [_int] * 8
        L.ldc_oracle_colconv.argtypes = [_f32p, _f32p, _f32p] + [_int] * 6
        L.ldc_oracle_bn_silu.argtypes = [_f32p] * 8 + [_int] * 4 + [_flt, _flt, _int]
        L.ldc_oracle_bn_silu_bwd.argtypes = [_f32p] * 9 + [_int] * 5
        L.ldc_oracle_colconv_bwd.argtypes = [_f32p] * 5 + [_int] * 6
        L.ldc_oracle_sample_bwd.argtypes = [_f32p, _f32p, _f32p, _i64p, _f64p, _f32p] + [_int] * 8
        L.ldc_oracle_offset_conv_bwd.argtypes = [_f32p, _f32p, _f32p, _f64p, _f32p, _f32p] + [_int] * 6
        L.ldc_oracle_num_threads.restype = _int
    return _lib


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def out_size(H: int, s: int) -> int:
    return (H - 1) // s + 1


def p_n(N: int) -> np.ndarray:
    """conv.py:413-432 -> int64 (1,2N,1,1), rows first then columns."""
    out = np.zeros(2 * N, dtype=np.int64)
    lib().ldc_oracle_p_n(N, _p(out, _i64p))
    return out.reshape(1, 2 * N, 1, 1)


def offset_conv(x, w, b, N: int, s: int) -> np.ndarray:
    x, w = _f32(x), _f32(w)
    b = None if b is None else _f32(b)
    B, C, H, W = x.shape
    h, wo = out_size(H, s), out_size(W, s)
    off = np.empty((B, 2 * N, h, wo), dtype=np.float32)
    lib().ldc_oracle_offset_conv(_p(x, _f32p), _p(w, _f32p), _p(b, _f32p), _p(off, _f32p), B, C, H, W, N, s)
    return off


def grid(off, H: int, W: int, N: int, s: int):
    """-> idx (B,h,w,N,4) int32 {r0,r1,k0,k1}, coord (B,h,w,N,2) {pcr,pck}, g (B,h,w,N,4) {lt,rb,lb,rt}."""
    off = _f32(off)
    B, _, h, w = off.shape
    pn = p_n(N).reshape(-1)
    idx = np.empty((B, h, w, N, 4), dtype=np.int32)
    coord = np.empty((B, h, w, N, 2), dtype=np.float32)
    g = np.empty((B, h, w, N, 4), dtype=np.float32)
    lib().ldc_oracle_grid(_p(off, _f32p), _p(pn, _i64p), _p(idx, _i32p), _p(coord, _f32p), _p(g, _f32p),
                          B, H, W, h, w, N, s)
    return idx, coord, g


def sample(x, off, N: int, s: int) -> np.ndarray:
    """-> x_offset (B,C,h*N,w), the tensor the reference feeds its (N,1) conv (conv.py:407)."""
    x, off = _f32(x), _f32(off)
    B, C, H, W = x.shape
    _, _, h, w = off.shape
    pn = p_n(N).reshape(-1)
    xo = np.empty((B, C, h * N, w), dtype=np.float32)
    lib().ldc_oracle_sample(_p(x, _f32p), _p(off, _f32p), _p(pn, _i64p), _p(xo, _f32p), B, C, H, W, h, w, N, s)
    return xo


def colconv(x_offset, wc, N: int) -> np.ndarray:
    x_offset, wc = _f32(x_offset), _f32(wc)
    B, C, hN, w = x_offset.shape
    h = hN // N
    O = wc.shape[0]
    pre = np.empty((B, O, h, w), dtype=np.float32)
    lib().ldc_oracle_colconv(_p(x_offset, _f32p), _p(wc, _f32p), _p(pre, _f32p), B, C, h, w, N, O)
    return pre


def bn_silu(pre, gamma, beta, running_mean, running_var, eps: float, momentum: float, training: bool):
    """-> out, save_mean, save_invstd; running_* (float32 arrays) are updated in place when training."""
    pre, gamma, beta = _f32(pre), _f32(gamma), _f32(beta)
    assert running_mean.dtype == np.float32 and running_var.dtype == np.float32
    B, O, h, w = pre.shape
    out = np.empty_like(pre)
    sm = np.empty(O, dtype=np.float32)
    si = np.empty(O, dtype=np.float32)
    lib().ldc_oracle_bn_silu(_p(pre, _f32p), _p(gamma, _f32p), _p(beta, _f32p), _p(running_mean, _f32p),
                             _p(running_var, _f32p), _p(sm, _f32p), _p(si, _f32p), _p(out, _f32p), B, O, h, w,
                             eps, momentum, int(training))
    return out, sm, si


def bn_silu_bwd(pre, grad_out, gamma, beta, mean, invstd, training: bool):
    pre, grad_out = _f32(pre), _f32(grad_out)
    gamma, beta, mean, invstd = _f32(gamma), _f32(beta), _f32(mean), _f32(invstd)
    B, O, h, w = pre.shape
    gp = np.empty_like(pre)
    gg = np.empty(O, dtype=np.float32)
    gb = np.empty(O, dtype=np.float32)
    lib().ldc_oracle_bn_silu_bwd(_p(pre, _f32p), _p(grad_out, _f32p), _p(gamma, _f32p), _p(beta, _f32p),
                                 _p(mean, _f32p), _p(invstd, _f32p), _p(gp, _f32p), _p(gg, _f32p), _p(gb, _f32p),
                                 B, O, h, w, int(training))
    return gp, gg, gb


def colconv_bwd(grad_pre, x_offset, wc, N: int):
    grad_pre, x_offset, wc = _f32(grad_pre), _f32(x_offset), _f32(wc)
    B, O, h, w = grad_pre.shape
    C = x_offset.shape[1]
    gxo = np.empty_like(x_offset)
    gw = np.empty((O, C, N, 1), dtype=np.float32)
    lib().ldc_oracle_colconv_bwd(_p(grad_pre, _f32p), _p(x_offset, _f32p), _p(wc, _f32p), _p(gxo, _f32p),
                                 _p(gw, _f32p), B, C, h, w, N, O)
    return gxo, gw


def sample_bwd(grad_x_offset, x, off, N: int, s: int):
    """-> grad_x (float64, B,C,H,W: scatter part only), grad_off (B,2N,h,w)."""
    g, x, off = _f32(grad_x_offset), _f32(x), _f32(off)
    B, C, H, W = x.shape
    _, _, h, w = off.shape
    pn = p_n(N).reshape(-1)
    gx = np.zeros((B, C, H, W), dtype=np.float64)
    goff = np.empty_like(off)
    lib().ldc_oracle_sample_bwd(_p(g, _f32p), _p(x, _f32p), _p(off, _f32p), _p(pn, _i64p), _p(gx, _f64p),
                                _p(goff, _f32p), B, C, H, W, h, w, N, s)
    return gx, goff


def offset_conv_bwd(grad_off, x, w, N: int, s: int, grad_x=None):
    """-> grad_x (float64; accumulated into `grad_x` when given), grad_w (2N,C,3,3), grad_b (2N)."""
    grad_off, x, w = _f32(grad_off), _f32(x), _f32(w)
    B, C, H, W = x.shape
    gx = np.zeros((B, C, H, W), dtype=np.float64) if grad_x is None else grad_x
    gw = np.empty((2 * N, C, 3, 3), dtype=np.float32)
    gb = np.empty(2 * N, dtype=np.float32)
    lib().ldc_oracle_offset_conv_bwd(_p(grad_off, _f32p), _p(x, _f32p), _p(w, _f32p), _p(gx, _f64p), _p(gw, _f32p),
                                     _p(gb, _f32p), B, C, H, W, N, s)
    return gx, gw, gb


@dataclass
class LDConvParams:
    """The reference LDConv state (conv.py:351-359), as numpy fp32."""
    p_conv_weight: np.ndarray   # (2N, C, 3, 3)
    p_conv_bias: np.ndarray     # (2N,)
    conv_weight: np.ndarray     # (O, C, N, 1), no bias
    bn_weight: np.ndarray       # (O,)
    bn_bias: np.ndarray         # (O,)
    running_mean: np.ndarray    # (O,)
    running_var: np.ndarray     # (O,)
    num_param: int
    stride: int
    eps: float = 1e-5
    momentum: float = 0.1


def forward(x, prm: LDConvParams, training: bool = False, offset=None, update_running: bool = True):
    """Whole LDConv.forward (conv.py:366-410).  Returns a dict with every intermediate the parity tests compare."""
    N, s = prm.num_param, prm.stride
    x = _f32(x)
    off = offset_conv(x, prm.p_conv_weight, prm.p_conv_bias, N, s) if offset is None else _f32(offset)
    xo = sample(x, off, N, s)
    pre = colconv(xo, prm.conv_weight, N)
    rm = prm.running_mean if update_running else prm.running_mean.copy()
    rv = prm.running_var if update_running else prm.running_var.copy()
    out, sm, si = bn_silu(pre, prm.bn_weight, prm.bn_bias, rm, rv, prm.eps, prm.momentum, training)
    return {"offset": off, "x_offset": xo, "pre": pre, "out": out, "save_mean": sm, "save_invstd": si}


def backward(x, prm: LDConvParams, fwd: dict, grad_out, training: bool):
    """Autograd of LDConv.forward (closed form).  Returns the gradients of x and of every
    parameter, keyed like the reference state_dict."""
    N, s = prm.num_param, prm.stride
    gp, gg, gb = bn_silu_bwd(fwd["pre"], grad_out, prm.bn_weight, prm.bn_bias, fwd["save_mean"], fwd["save_invstd"],
                             training)
    gxo, gwc = colconv_bwd(gp, fwd["x_offset"], prm.conv_weight, N)
    gx, goff = sample_bwd(gxo, x, fwd["offset"], N, s)
    gx, gpw, gpb = offset_conv_bwd(goff, x, prm.p_conv_weight, N, s, grad_x=gx)
    return {"x": gx.astype(np.float32), "offset": goff, "pre": gp, "x_offset": gxo, "conv.0.weight": gwc,
            "conv.1.weight": gg, "conv.1.bias": gb, "p_conv.weight": gpw, "p_conv.bias": gpb}
