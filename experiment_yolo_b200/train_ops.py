"""Training-graph glue around LDConv that runs through the library instead of ATen (SURVEY.md 8f "next" rows): ops with autograd
whose forward AND backward are one C-ABI call each.  They exist because the eager versions dominate the config-4 step by dtype
round trips, not by arithmetic: under bf16 autocast `upsample_nearest2d` ran in fp32 (the Concat behind it was promoted to fp32 and
cast back for the next conv), 4.5 ms of a 96 ms step for two YAML rows and the two SSFF interpolations.
"""
from __future__ import annotations

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def nhwc_view(t: torch.Tensor):
    """(B,C,H,W) tensor -> (NHWC view, pixel stride) without a copy when the channels are unit-stride and the pixels are evenly
    spaced (a dense channels_last tensor or a channel slice of one); otherwise a dense NHWC copy."""
    B, C, H, W = t.shape
    v = t.permute(0, 2, 3, 1)
    ld = v.stride(2)
    if v.stride(3) == 1 and ld >= C and ld % 8 == 0 and v.stride(1) == W * ld and v.stride(0) == H * W * ld and v.data_ptr() % 16 == 0:
        return v, ld
    v = v.contiguous()
    return v, C


class _UpsampleNearestFunction(torch.autograd.Function):
    """nn.Upsample(scale_factor=f, mode="nearest") for a bf16 NHWC tensor (yolov8-LD-P2.yaml:26,33; SSFF's interpolations,
    nn/extra_modules/block.py:3436-3437): forward = ldconv_upsample_nearest, backward = ldconv_upsample_nearest_bwd."""

    @staticmethod
    def forward(ctx, x, f):
        B, C, H, W = x.shape
        xv, ldx = nhwc_view(x)
        out = torch.empty((B, H * f, W * f, C), device=x.device, dtype=x.dtype)
        _lib.check(_lib.load().ldconv_upsample_nearest(xv.data_ptr(), ldx, out.data_ptr(), C, B, H, W, C, f, _lib.BF16, _stream()),
                   "ldconv_upsample_nearest")
        ctx.f = f
        return out.permute(0, 3, 1, 2)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        f = ctx.f
        B, C, Ho, Wo = grad_out.shape
        H, W = Ho // f, Wo // f
        gv, ldg = nhwc_view(grad_out if grad_out.dtype == torch.bfloat16 else grad_out.to(torch.bfloat16))
        gx = torch.empty((B, H, W, C), device=grad_out.device, dtype=torch.bfloat16)
        _lib.check(_lib.load().ldconv_upsample_nearest_bwd(gv.data_ptr(), ldg, gx.data_ptr(), C, B, H, W, C, f, _lib.BF16, _stream()),
                   "ldconv_upsample_nearest_bwd")
        return gx.permute(0, 3, 1, 2), None


def upsample_nearest(x: torch.Tensor, factor: int):
    """Nearest up-sampling by an integer factor through the library, or None when the case is not covered (the caller then runs
    torch's op): needs a CUDA bf16 4-D tensor with C % 8 == 0."""
    if not (x.is_cuda and x.dim() == 4 and x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0 and factor >= 1 and x.numel() > 0):
        return None
    return _UpsampleNearestFunction.apply(x, int(factor))


class _AddFunction(torch.autograd.Function):
    """`Add` row of the YAML (nn/extra_modules/block.py:3479-3484: torch.sum(torch.stack(x), 0)) for bf16 NHWC tensors: one
    ldconv_add_nhwc launch per four inputs (fp32 sum, one rounding) instead of stack -> sum -> layout / dtype copies; the gradient
    of a sum is the incoming gradient for every input."""

    @staticmethod
    def forward(ctx, *xs):
        import ctypes
        B, C, H, W = xs[0].shape
        pending = [nhwc_view(t) for t in xs]
        L = _lib.load()
        while True:
            part, pending = pending[:4], pending[4:]
            srcs = (ctypes.c_void_p * len(part))(*[v.data_ptr() for v, _ in part])
            lds = (ctypes.c_int * len(part))(*[ld for _, ld in part])
            out = torch.empty((B, H, W, C), device=xs[0].device, dtype=torch.bfloat16)
            _lib.check(L.ldconv_add_nhwc(srcs, lds, len(part), out.data_ptr(), C, B * H * W, C, _lib.BF16, _stream()), "ldconv_add_nhwc")
            if not pending:
                break
            pending.insert(0, (out, C))
        ctx.n = len(xs)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        return (grad_out,) * ctx.n


def add_maps(xs):
    """Sum of same-shape CUDA maps through the library, or None when the case is not covered: bf16 4-D tensors with C % 8 == 0
    (fp32 inputs are accepted under bf16 autocast and rounded to bf16 first -- the convs behind the row would do that anyway)."""
    xs = list(xs)
    if not xs or not all(t.is_cuda and t.dim() == 4 and t.shape == xs[0].shape for t in xs) or xs[0].shape[1] % 8 != 0 or xs[0].numel() == 0:
        return None
    ac = torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
    if not all(t.dtype == torch.bfloat16 or (ac and t.dtype == torch.float32) for t in xs):
        return None
    return _AddFunction.apply(*[t if t.dtype == torch.bfloat16 else t.to(torch.bfloat16) for t in xs])
