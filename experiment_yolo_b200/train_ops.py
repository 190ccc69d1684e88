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


class _ScalSeqTailFunction(torch.autograd.Function):
    """SSFF after its three 1x1 Conv blocks (nn/extra_modules/block.py:3432-3443), training mode, bf16: nearest up-sampling of the two
    coarser maps to the finest size, torch.stack along a depth axis, Conv3d(1x1x1, bias), BatchNorm3d (batch statistics),
    LeakyReLU(0.1), MaxPool3d((3,1,1)).  Through the library: the three slices are the row blocks of one (3 M, C) NHWC matrix (the
    up-sampling kernel writes the coarser two straight into their blocks), the Conv3d is the tcgen05 GEMM with the BatchNorm sums in
    its epilogue, ldconv_ssff_max_fwd applies BatchNorm + LeakyReLU + the depth maximum in one pass; backward = ldconv_ssff_max_bwd
    (routing to the arg-max slice) + the BatchNorm backward passes + the data / weight GEMMs + the up-sampling backward.  The eager
    version ran BatchNorm3d and its backward in fp32 on the 5-D volume and MaxPool3d through the generic 3-D pooling kernels:
    ~10 ms of an 85 ms step.  The Conv3d bias only moves the BatchNorm mean: its gradient is zero, the running mean includes it."""

    @staticmethod
    def forward(ctx, fine, mid, coarse, w3d, b3d, gamma, beta, running_mean, running_var, eps, momentum):
        L = _lib.load()
        st = _stream()
        B, C, H, W = fine.shape
        M = B * H * W
        dev = fine.device
        a = torch.empty((3 * M, C), device=dev, dtype=torch.bfloat16)
        factors = []
        for d, t in enumerate((fine, mid, coarse)):
            f = H // t.shape[2]
            factors.append(f)
            tv, ld = nhwc_view(t)
            _lib.check(L.ldconv_upsample_nearest(tv.data_ptr(), ld, a[d * M:].data_ptr(), C, B, t.shape[2], t.shape[3], C, f, _lib.BF16, st),
                       "ldconv_upsample_nearest")
        wt = w3d.detach().reshape(C, C).to(torch.bfloat16).contiguous()                 # (O, K)
        stats = torch.zeros((2, C), device=dev, dtype=torch.float64)
        pre = torch.empty((3 * M, C), device=dev, dtype=torch.bfloat16)
        _lib.check(L.ldconv_gemm_fwd(a.data_ptr(), wt.data_ptr(), None, None, None, pre.data_ptr(), stats[0].data_ptr(), stats[1].data_ptr(),
                                     3 * M, C, C, _lib.ACT_NONE, _lib.BF16, st), "ldconv_gemm_fwd")
        scale = torch.empty(C, device=dev, dtype=torch.float32)
        shift = torch.empty(C, device=dev, dtype=torch.float32)
        mean = torch.empty(C, device=dev, dtype=torch.float32)
        invstd = torch.empty(C, device=dev, dtype=torch.float32)
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        bias = None if b3d is None else b3d.detach().float()
        if bias is not None:
            running_mean.sub_(bias)          # BatchNorm sees acc + bias: only the running mean notices
        _lib.check(L.ldconv_bn_finalize(stats[0].data_ptr(), stats[1].data_ptr(), 3 * M, g32.data_ptr(), b32.data_ptr(),
                                        running_mean.data_ptr(), running_var.data_ptr(), float(eps), float(momentum), 1, scale.data_ptr(),
                                        shift.data_ptr(), mean.data_ptr(), invstd.data_ptr(), C, st), "ldconv_bn_finalize")
        if bias is not None:
            running_mean.add_(bias)
        out = torch.empty((B, H, W, C), device=dev, dtype=torch.bfloat16)
        _lib.check(L.ldconv_ssff_max_fwd(pre.data_ptr(), scale.data_ptr(), shift.data_ptr(), out.data_ptr(), M, C, _lib.BF16, st),
                   "ldconv_ssff_max_fwd")
        ctx.save_for_backward(a, pre, scale, shift, mean, invstd, wt)
        ctx.meta = (B, C, H, W, factors, [t.shape for t in (fine, mid, coarse)], w3d.shape, w3d.dtype, gamma.dtype, b3d is not None,
                    None if b3d is None else b3d.dtype)
        return out.permute(0, 3, 1, 2)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        L = _lib.load()
        st = _stream()
        a, pre, scale, shift, mean, invstd, wt = ctx.saved_tensors
        B, C, H, W, factors, shapes, w_shape, w_dtype, g_dtype, has_bias, b_dtype = ctx.meta
        M = B * H * W
        dev = a.device
        go = grad_out if grad_out.dtype == torch.bfloat16 else grad_out.to(torch.bfloat16)
        gv, ld = nhwc_view(go)
        if ld != C:
            gv = gv.contiguous()
        dz = torch.empty_like(pre)
        _lib.check(L.ldconv_ssff_max_bwd(pre.data_ptr(), scale.data_ptr(), shift.data_ptr(), gv.data_ptr(), dz.data_ptr(), M, C, _lib.BF16, st),
                   "ldconv_ssff_max_bwd")
        red = torch.zeros((2, C), device=dev, dtype=torch.float64)
        _lib.check(L.ldconv_bn_act_bwd_reduce(pre.data_ptr(), dz.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                              invstd.data_ptr(), red.data_ptr(), 3 * M, C, _lib.ACT_NONE, _lib.BF16, st),
                   "ldconv_bn_act_bwd_reduce")
        gpre = torch.empty_like(pre)
        _lib.check(L.ldconv_bn_act_bwd_apply(pre.data_ptr(), dz.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                                             invstd.data_ptr(), red.data_ptr(), gpre.data_ptr(), 3 * M, C, _lib.ACT_NONE, 1, _lib.BF16, st),
                   "ldconv_bn_act_bwd_apply")
        grad_w = torch.zeros((C, C), device=dev, dtype=torch.float32)
        _lib.check(L.ldconv_gemm_bwd_weight(gpre.data_ptr(), a.data_ptr(), grad_w.data_ptr(), 3 * M, C, C, _lib.BF16, st),
                   "ldconv_gemm_bwd_weight")
        wt_t = wt.t().contiguous()
        ga = dz                              # reuse: the routed gradient is dead once grad_pre exists
        _lib.check(L.ldconv_gemm_fwd(gpre.data_ptr(), wt_t.data_ptr(), None, None, ga.data_ptr(), None, None, None, 3 * M, C, C,
                                     _lib.ACT_NONE, _lib.BF16, st), "ldconv_gemm_fwd(data grad)")
        grads = []
        for d in range(3):
            f = factors[d]
            blk = ga[d * M:(d + 1) * M]
            if f == 1:
                grads.append(blk.view(B, H, W, C).permute(0, 3, 1, 2))
            else:
                h, w = shapes[d][2], shapes[d][3]
                gx = torch.empty((B, h, w, C), device=dev, dtype=torch.bfloat16)
                _lib.check(L.ldconv_upsample_nearest_bwd(blk.data_ptr(), C, gx.data_ptr(), C, B, h, w, C, f, _lib.BF16, st),
                           "ldconv_upsample_nearest_bwd")
                grads.append(gx.permute(0, 3, 1, 2))
        grad_b = torch.zeros(C, device=dev, dtype=b_dtype) if has_bias else None
        return (grads[0], grads[1], grads[2], grad_w.view(w_shape).to(w_dtype), grad_b, red[1].to(g_dtype), red[0].to(g_dtype),
                None, None, None, None)


def scalseq_tail(fine, mid, coarse, conv3d, bn):
    """The SSFF tail in training mode through the library (see _ScalSeqTailFunction), or None when the case is not covered: CUDA bf16
    maps with C % 8 == 0 whose sizes divide the finest one by an integer, a 1x1x1 Conv3d C -> C, affine BatchNorm3d with fp32 running
    statistics."""
    ts = (fine, mid, coarse)
    B, C, H, W = fine.shape
    if not all(t.is_cuda and t.dim() == 4 and t.dtype == torch.bfloat16 and t.shape[0] == B and t.shape[1] == C for t in ts):
        return None
    if C % 8 != 0 or C > 256 or fine.numel() == 0 or 3 * B * H * W >= 2 ** 31:
        return None
    for t in (mid, coarse):
        h, w = t.shape[2:]
        if h == 0 or w == 0 or H % h != 0 or W % w != 0 or H // h != W // w:
            return None
    if tuple(conv3d.weight.shape) != (C, C, 1, 1, 1) or not (bn.training and bn.affine and bn.track_running_stats and
                                                             bn.running_mean is not None and bn.running_mean.dtype == torch.float32):
        return None
    momentum = bn.momentum
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
    return _ScalSeqTailFunction.apply(fine, mid, coarse, conv3d.weight, conv3d.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                      bn.eps, momentum)
