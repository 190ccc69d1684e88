"""Drop-in `LDConv` (Linear Deformable Convolution) whose forward and backward run as sm_100a CUDA kernels.

Host-side mirror of the reference module /root/reference/ultralytics/nn/modules/conv.py:350-503:
same constructor `(inc, outc, num_param, stride=1, bias=None)`, same sub-modules and registration order
(`conv = Sequential(Conv2d((N,1),(N,1)), BatchNorm2d, SiLU)`, `p_conv = Conv2d(3x3)`, int64 buffer `p_n`), hence the same
`state_dict()` keys / shapes / dtypes and the same random initialisation under the same seed, and the same YAML row
`[-1, 1, LDConv, [c2, num_param, stride]]` (nn/tasks.py:813-864).  PyTorch is used for device memory and streams only;
every op of the path goes through the C ABI of libldconv_b200.so (include/ldconv_b200.h).  No CPU fallback: a CPU tensor
raises RuntimeError (its text contains 'CUDA tensor', which is what the reference's stride probe at nn/tasks.py:317-321
looks for before retrying on the GPU).

Numerics contract: sampling offsets, coordinates and bilinear weights are always fp32, also for bf16
activations -- the reference's own low-precision coordinate math is broken.  bf16 parity is defined
against the fp32 reference evaluated on bf16-rounded inputs and parameters.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import _lib

__all__ = ["LDConv", "install", "ldconv_function", "base_grid"]

_DTYPES = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}
# float16 (the reference's only reduced-precision mode: `model.half()` in get_FPS.py:59-61 / engine/validator.py:113-115 and
# fp16 autocast in engine/trainer.py:800) is accepted at the module boundary and computed by the fp32 kernels: fp16 -> fp32 is
# exact, the result is rounded to fp16 once on the way out, gradients come back in the parameter dtype.  Parity: the fp32
# reference on fp16-rounded tensors (tests/test_gpu_parity.py::test_module_half_*).
_COMPUTE_AS = {torch.float16: torch.float32}


def _ver(t: Optional[torch.Tensor]):
    """cache-key component of a tensor: (address, version counter).  Inference tensors have no version counter
    (`_version` raises): they are keyed by address and a marker, and never shared with a grad-mode forward (see
    LDConv._prepared)."""
    if t is None:
        return None
    return (t.data_ptr(), -1 if t.is_inference() else t._version)


def base_grid(num_param: int) -> torch.Tensor:
    """conv.py:413-432 (`_get_p_n`): the un-centred raster base grid as the reference's int64 (1,2N,1,1) buffer,
    rows (H axis) first, then columns."""
    base = round(math.sqrt(num_param))
    rows = [i // base for i in range(num_param)]
    cols = [i % base for i in range(num_param)]
    return torch.tensor(rows + cols, dtype=torch.int64).view(1, 2 * num_param, 1, 1)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check_input(x: torch.Tensor) -> int:
    if not x.is_cuda:
        raise RuntimeError(
            "experiment_yolo_b200.LDConv runs on sm_100a only and needs a CUDA tensor "
            f"(got a {x.device.type} tensor); there is no CPU fallback")
    if x.dtype not in _DTYPES:
        raise TypeError(f"experiment_yolo_b200.LDConv supports float32, bfloat16 and float16 activations, got {x.dtype}")
    if x.dim() != 4:
        raise ValueError(f"LDConv expects a (B,C,H,W) tensor, got shape {tuple(x.shape)}")
    return _DTYPES[x.dtype]


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) logical -> dense (B,H,W,C) view; zero-copy when x is already channels_last."""
    return x.permute(0, 2, 3, 1).contiguous()


class _Prepared:
    """Per-call operand forms of the parameters (tiny tensors; cached by the module while the parameters are unchanged)."""
    __slots__ = ("w_off", "b_off", "w_off_tc", "w_off_s2d", "wt", "wt_t", "pn", "key")


def _prepare(p_w, p_b, c_w, p_n, dtype: torch.dtype, need_wt_t: bool) -> _Prepared:
    pr = _Prepared()
    N = c_w.shape[2]
    O, C = c_w.shape[0], c_w.shape[1]
    # offset conv weights: (2N,C,3,3) -> (3,3,C,2N) fp32, rounded through the activation dtype first so that the bf16
    # path sees exactly the bf16-rounded parameters the parity oracle uses
    pr.w_off = p_w.detach().to(dtype).float().permute(2, 3, 1, 0).contiguous()
    pr.b_off = None if p_b is None else p_b.detach().to(dtype).float().contiguous()
    # tensor-core offset conv (bf16): (2N,C,3,3) -> (2N, 3,3,C) -> (2N, 9C), k = tap*C + c
    pr.w_off_tc = (p_w.detach().to(dtype).permute(0, 2, 3, 1).reshape(2 * N, 9 * C).contiguous()
                   if dtype == torch.bfloat16 else None)
    # stride-2 offset conv on the space-to-depth view (ldconv_offset_conv_s2d_fwd): (2N,C,3,3) -> (2N, ty, tx, sy, sx, C),
    # ky = 0,1,2 -> (ty, sy) = (0,1), (1,0), (1,1); the unused (ty=0, sy=0) / (tx=0, sx=0) combinations stay zero
    pr.w_off_s2d = None
    if dtype == torch.bfloat16 and C in (16, 32, 64) and 2 * N <= 16:
        w6 = torch.zeros((2 * N, 2, 2, 2, 2, C), device=p_w.device, dtype=dtype)
        tmap = ((0, 1), (1, 0), (1, 1))
        pw = p_w.detach().to(dtype)
        for ky in range(3):
            for kx in range(3):
                (ty, sy), (tx, sx) = tmap[ky], tmap[kx]
                w6[:, ty, tx, sy, sx, :] = pw[:, :, ky, kx]
        pr.w_off_s2d = w6.reshape(2 * N, 16 * C).contiguous()
    # (N,1) conv weight (O,C,N,1) -> (O, N*C), k = n*C + c: the order the NHWC gather produces
    pr.wt = c_w.detach().to(dtype).reshape(O, C, N).permute(0, 2, 1).reshape(O, N * C).contiguous()
    pr.wt_t = pr.wt.t().contiguous() if need_wt_t else None
    pr.pn = p_n.detach().reshape(-1).to(device=c_w.device, dtype=torch.int32).contiguous()
    return pr


def offset_conv_nhwc(xh: torch.Tensor, pr: _Prepared, N: int, s: int) -> torch.Tensor:
    """conv.py:356,368 `offset = p_conv(x)` on a dense NHWC tensor -> (B,h,w,2N) fp32 offsets, rows then columns.  Picks the
    kernel the shape allows: zero-copy tcgen05 on the space-to-depth view (stride 2), tcgen05 implicit GEMM, CUDA cores."""
    L = _lib.load()
    B, H, W, C = xh.shape
    dt = _DTYPES[xh.dtype]
    st = _stream()
    h, w = (H - 1) // s + 1, (W - 1) // s + 1
    off = torch.empty((B, h, w, 2 * N), device=xh.device, dtype=torch.float32)
    if s == 2 and pr.w_off_s2d is not None and L.ldconv_offset_conv_s2d_supported(C, N, H, W, dt):
        _lib.check(L.ldconv_offset_conv_s2d_fwd(_ptr(xh), _ptr(pr.w_off_s2d), _ptr(pr.b_off), _ptr(off), B, C, H, W, N, dt, st),
                   "ldconv_offset_conv_s2d_fwd")
    elif pr.w_off_tc is not None and L.ldconv_offset_conv_tc_supported(C, N, s, dt):
        _lib.check(L.ldconv_offset_conv_tc_fwd(_ptr(xh), _ptr(pr.w_off_tc), _ptr(pr.b_off), _ptr(off), B, C, H, W, N, s, dt, st),
                   "ldconv_offset_conv_tc_fwd")
    else:
        _lib.check(L.ldconv_offset_conv_fwd(_ptr(xh), _ptr(pr.w_off), _ptr(pr.b_off), _ptr(off), B, C, H, W, N, s, dt, st),
                   "ldconv_offset_conv_fwd")
    return off


class _LDConvFunction(torch.autograd.Function):
    """Forward = offset conv -> fused grid+gather -> GEMM (+BN statistics) -> BN/SiLU; backward = the closed form of
    SURVEY.md Appendix A.  Every step is one C-ABI call."""

    # bf16 activations: accumulate grad_x in bf16 (ldconv_gather_bwd_acc16); False = the fp32 accumulator + cast pass (A/B, tests)
    bf16_accumulator = True

    @staticmethod
    def forward(ctx, x, p_w, p_b, c_w, c_b, bn_w, bn_b, running_mean, running_var, p_n, stride, eps, momentum, training,
                prepared):
        L = _lib.load()
        dt = _check_input(x)
        st = _stream()
        B, C, H, W = x.shape
        O, N = c_w.shape[0], c_w.shape[2]
        s = int(stride)
        h, w = (H - 1) // s + 1, (W - 1) // s + 1
        M, K = B * h * w, N * C
        dev = x.device
        needs_grad = any(ctx.needs_input_grad)
        pr = prepared if prepared is not None else _prepare(p_w, p_b, c_w, p_n, x.dtype, needs_grad)
        if needs_grad and pr.wt_t is None:
            pr.wt_t = pr.wt.t().contiguous()

        xh = _nhwc(x)
        off = offset_conv_nhwc(xh, pr, N, s)
        operand = torch.empty((M, K), device=dev, dtype=x.dtype)
        _lib.check(L.ldconv_gather_fwd(_ptr(xh), _ptr(off), _ptr(pr.pn), _ptr(operand), None, None, B, C, H, W, N, s, dt, st),
                   "ldconv_gather_fwd")

        gamma = None if bn_w is None else bn_w.detach().float().contiguous()
        beta = None if bn_b is None else bn_b.detach().float().contiguous()
        scale = torch.empty(O, device=dev, dtype=torch.float32)
        shift = torch.empty(O, device=dev, dtype=torch.float32)
        save_mean = torch.empty(O, device=dev, dtype=torch.float32)
        save_invstd = torch.empty(O, device=dev, dtype=torch.float32)
        out = torch.empty((B, h, w, O), device=dev, dtype=x.dtype)
        batch_stats = bool(training) or running_mean is None
        cbias = None if c_b is None else c_b.detach().float()
        pre = None
        if batch_stats:
            if M <= 1:
                # torch.nn.functional.batch_norm raises the same error for the reference module
                raise ValueError(f"Expected more than 1 value per channel when training, got input size {(B, O, h, w)}")
            stats = torch.zeros((2, O), device=dev, dtype=torch.float64)
            pre = torch.empty((M, O), device=dev, dtype=x.dtype)
            _lib.check(L.ldconv_gemm_fwd(_ptr(operand), _ptr(pr.wt), None, None, None, _ptr(pre), _ptr(stats[0]),
                                         _ptr(stats[1]), M, K, O, _lib.ACT_NONE, dt, st), "ldconv_gemm_fwd")
            rm32 = rv32 = None
            if training and running_mean is not None:
                rm32 = running_mean if running_mean.dtype == torch.float32 else running_mean.float()
                rv32 = running_var if running_var.dtype == torch.float32 else running_var.float()
                if cbias is not None:      # BN sees acc + bias: only the running mean notices
                    rm32.sub_(cbias)
            _lib.check(L.ldconv_bn_finalize(_ptr(stats[0]), _ptr(stats[1]), M, _ptr(gamma), _ptr(beta), _ptr(rm32),
                                            _ptr(rv32), float(eps), float(momentum), 1, _ptr(scale), _ptr(shift),
                                            _ptr(save_mean), _ptr(save_invstd), O, st), "ldconv_bn_finalize")
            if rm32 is not None:
                if cbias is not None:
                    rm32.add_(cbias)
                if rm32 is not running_mean:
                    running_mean.copy_(rm32)
                    running_var.copy_(rv32)
            _lib.check(L.ldconv_bn_act_apply(_ptr(pre), _ptr(scale), _ptr(shift), _ptr(out), M, O, _lib.ACT_SILU, dt, st),
                       "ldconv_bn_act_apply")
        else:
            rm32 = running_mean.float() if (running_mean.dtype != torch.float32 or cbias is not None) else running_mean
            if cbias is not None:
                rm32 = rm32 - cbias
            rv32 = running_var if running_var.dtype == torch.float32 else running_var.float()
            _lib.check(L.ldconv_bn_finalize(None, None, 0, _ptr(gamma), _ptr(beta), _ptr(rm32), _ptr(rv32), float(eps),
                                            0.0, 0, _ptr(scale), _ptr(shift), _ptr(save_mean), _ptr(save_invstd), O, st),
                       "ldconv_bn_finalize")
            if needs_grad:
                pre = torch.empty((M, O), device=dev, dtype=x.dtype)
            _lib.check(L.ldconv_gemm_fwd(_ptr(operand), _ptr(pr.wt), _ptr(scale), _ptr(shift), _ptr(out), _ptr(pre), None,
                                         None, M, K, O, _lib.ACT_SILU, dt, st), "ldconv_gemm_fwd")
        if needs_grad:
            ctx.save_for_backward(xh, off, operand, pre, scale, shift, save_mean, save_invstd, pr.wt_t, pr.w_off, pr.pn)
            ctx.meta = (B, C, H, W, O, N, s, dt, batch_stats, x.dtype, p_w.dtype, c_w.dtype,
                        None if bn_w is None else bn_w.dtype, c_b is not None)
        return out.permute(0, 3, 1, 2)   # logical NCHW, channels_last strides: downstream convs / cats run unchanged

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        L = _lib.load()
        xh, off, operand, pre, scale, shift, save_mean, save_invstd, wt_t, w_off, pn = ctx.saved_tensors
        B, C, H, W, O, N, s, dt, batch_stats, xdtype, pw_dtype, cw_dtype, bn_dtype, has_cbias = ctx.meta
        st = _stream()
        dev = xh.device
        h, w = (H - 1) // s + 1, (W - 1) // s + 1
        M, K = B * h * w, N * C
        need_x = ctx.needs_input_grad[0]

        go = _nhwc(grad_out.to(xdtype))
        red = torch.zeros((2, O), device=dev, dtype=torch.float64)
        _lib.check(L.ldconv_bn_act_bwd_reduce(_ptr(pre), _ptr(go), _ptr(scale), _ptr(shift), _ptr(save_mean),
                                              _ptr(save_invstd), _ptr(red), M, O, _lib.ACT_SILU, dt, st),
                   "ldconv_bn_act_bwd_reduce")
        grad_pre = torch.empty((M, O), device=dev, dtype=xdtype)
        _lib.check(L.ldconv_bn_act_bwd_apply(_ptr(pre), _ptr(go), _ptr(scale), _ptr(shift), _ptr(save_mean),
                                             _ptr(save_invstd), _ptr(red), _ptr(grad_pre), M, O, _lib.ACT_SILU,
                                             int(batch_stats), dt, st), "ldconv_bn_act_bwd_apply")
        # weight gradient of the (N,1) conv: (O,K) fp32, k = n*C + c  ->  (O,C,N,1)
        grad_wt = torch.zeros((O, K), device=dev, dtype=torch.float32)
        _lib.check(L.ldconv_gemm_bwd_weight(_ptr(grad_pre), _ptr(operand), _ptr(grad_wt), M, K, O, dt, st),
                   "ldconv_gemm_bwd_weight")
        grad_c_w = grad_wt.view(O, N, C).permute(0, 2, 1).reshape(O, C, N, 1).to(cw_dtype)
        # data gradient: grad_operand (M,K) = grad_pre (M,O) . W (O,K)   (same GEMM kernel, roles of K and O swapped)
        grad_operand = torch.empty((M, K), device=dev, dtype=xdtype)
        _lib.check(L.ldconv_gemm_fwd(_ptr(grad_pre), _ptr(wt_t), None, None, _ptr(grad_operand), None, None, None, M, O, K,
                                     _lib.ACT_NONE, dt, st), "ldconv_gemm_fwd(data grad)")
        # grad_x accumulator: bf16 for the bf16 path (eight channels per reduction request instead of four, no fp32 buffer and
        # no cast pass; include/ldconv_b200.h states its tolerance), fp32 otherwise
        acc16 = bool(need_x and dt == _lib.BF16 and _LDConvFunction.bf16_accumulator
                     and L.ldconv_bwd_acc16_supported(B, C, H, W, N, s))
        grad_off = torch.empty((B, h, w, 2 * N), device=dev, dtype=torch.float32)
        grad_w_off = torch.zeros((3, 3, C, 2 * N), device=dev, dtype=torch.float32)
        grad_b_off = torch.zeros((2 * N,), device=dev, dtype=torch.float32)
        ws_bytes = int(L.ldconv_offset_conv_bwd_workspace_bytes(B, C, H, W, N, s, dt)) if dt == _lib.BF16 else 0
        if acc16:
            grad_xa = torch.zeros((B, H, W, C), device=dev, dtype=torch.bfloat16)
            _lib.check(L.ldconv_gather_bwd_acc16(_ptr(grad_operand), _ptr(xh), _ptr(off), _ptr(pn), _ptr(grad_xa), _ptr(grad_off),
                                                 B, C, H, W, N, s, st), "ldconv_gather_bwd_acc16")
            ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
            _lib.check(L.ldconv_offset_conv_bwd_tc_acc16(_ptr(grad_off), _ptr(xh), _ptr(w_off), _ptr(grad_xa), _ptr(grad_w_off),
                                                         _ptr(grad_b_off), _ptr(ws), ws_bytes, B, C, H, W, N, s, st),
                       "ldconv_offset_conv_bwd_tc_acc16")
            grad_x = grad_xa.permute(0, 3, 1, 2)
        else:
            grad_x32 = torch.zeros((B, H, W, C), device=dev, dtype=torch.float32) if need_x else None
            _lib.check(L.ldconv_gather_bwd(_ptr(grad_operand), _ptr(xh), _ptr(off), _ptr(pn), _ptr(grad_x32), _ptr(grad_off),
                                           B, C, H, W, N, s, dt, st), "ldconv_gather_bwd")
            if ws_bytes > 0:      # bf16: weight gradient as a tensor-core reduction over an L2-resident im2col workspace
                ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
                _lib.check(L.ldconv_offset_conv_bwd_tc(_ptr(grad_off), _ptr(xh), _ptr(w_off), _ptr(grad_x32), _ptr(grad_w_off),
                                                       _ptr(grad_b_off), _ptr(ws), ws_bytes, B, C, H, W, N, s, dt, st),
                           "ldconv_offset_conv_bwd_tc")
            else:
                _lib.check(L.ldconv_offset_conv_bwd(_ptr(grad_off), _ptr(xh), _ptr(w_off), _ptr(grad_x32), _ptr(grad_w_off),
                                                    _ptr(grad_b_off), B, C, H, W, N, s, dt, st), "ldconv_offset_conv_bwd")
            grad_x = grad_x32.to(xdtype).permute(0, 3, 1, 2) if need_x else None
        grad_p_w = grad_w_off.permute(3, 2, 0, 1).contiguous().to(pw_dtype)
        grad_p_b = grad_b_off.to(pw_dtype)
        grad_gamma = red[1].to(bn_dtype) if bn_dtype is not None else None
        grad_beta = red[0].to(bn_dtype) if bn_dtype is not None else None
        grad_c_b = grad_pre.float().sum(0).to(cw_dtype) if has_cbias else None
        return (grad_x, grad_p_w, grad_p_b, grad_c_w, grad_c_b, grad_gamma, grad_beta, None, None, None, None, None,
                None, None, None)


def ldconv_function(x, p_conv_weight, p_conv_bias, conv_weight, bn_weight, bn_bias, running_mean, running_var, p_n,
                    stride: int, eps: float = 1e-5, momentum: float = 0.1, training: bool = False, conv_bias=None,
                    prepared=None):
    """Functional form of the whole module (parameters in the reference's own layouts)."""
    return _LDConvFunction.apply(x, p_conv_weight, p_conv_bias, conv_weight, conv_bias, bn_weight, bn_bias, running_mean,
                                 running_var, p_n, stride, eps, momentum, training, prepared)


def ldconv_fused_inference(x, prepared: _Prepared, scale, shift, C, O, N, s):
    """One-kernel bf16 inference forward (ldconv_fused_fwd): the resampled operand never reaches HBM."""
    L = _lib.load()
    dt = _check_input(x)
    B, _, H, W = x.shape
    h, w = (H - 1) // s + 1, (W - 1) // s + 1
    xh = _nhwc(x)
    out = torch.empty((B, h, w, O), device=x.device, dtype=x.dtype)
    _lib.check(L.ldconv_fused_fwd(_ptr(xh), _ptr(prepared.w_off), _ptr(prepared.b_off), _ptr(prepared.pn),
                                  _ptr(prepared.wt), _ptr(scale), _ptr(shift), _ptr(out), None, B, C, H, W, N, s, O,
                                  _lib.ACT_SILU, dt, _stream()), "ldconv_fused_fwd")
    return out.permute(0, 3, 1, 2)


class LDConv(nn.Module):
    """B200-native LDConv.  Signature, children, buffer and state_dict layout: conv.py:350-359."""

    # small-C one-kernel inference path (the model's first row, C <= 4: ldconv_fused_fwd); class-level switch so tests can A/B it
    use_fused_inference = True
    # inference: gather + GEMM + BN + SiLU as one persistent kernel after the tensor-core offset conv (ldconv_gather_gemm_fwd)
    use_gather_gemm = True
    # inference: the whole forward in one kernel, offset conv on the tensor cores over the gather's own staged tile
    # (ldconv_onepass_fwd; the yolov8-LD-P2 shapes).  A/B switch for the tests, which compare the two paths bit for bit
    use_onepass = True

    def __init__(self, inc, outc, num_param, stride=1, bias=None):
        super().__init__()
        self.num_param = num_param
        self.stride = stride
        # same construction order as the reference => same parameter values under the same torch seed
        self.conv = nn.Sequential(
            nn.Conv2d(inc, outc, kernel_size=(num_param, 1), stride=(num_param, 1), bias=bias),
            nn.BatchNorm2d(outc),
            nn.SiLU())
        self.p_conv = nn.Conv2d(inc, 2 * num_param, kernel_size=3, padding=1, stride=stride)
        nn.init.constant_(self.p_conv.weight, 0)      # conv.py:357
        # conv.py:358,361-364: the reference registers a backward hook that scales nothing (it builds two generators and
        # returns None), so gradients are NOT multiplied by 0.1; reproduced by not registering anything.
        self.register_buffer("p_n", base_grid(num_param))
        self._prep_cache = None

    # ---- parameter-derived operand cache (not part of the state; dropped on pickle / deepcopy) --------------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_prep_cache"] = None
        return state

    def _prepared(self, dtype: torch.dtype, need_wt_t: bool):
        conv, pconv = self.conv[0], self.p_conv
        # torch.is_inference_mode_enabled(): operands built under inference_mode are inference tensors and cannot be saved for
        # backward, so a grad-mode forward never reuses them (the reference's smart_inference_mode validates between epochs)
        key = (dtype, _ver(conv.weight), _ver(pconv.weight), _ver(pconv.bias), _ver(self.p_n), conv.weight.device,
               torch.is_inference_mode_enabled())
        c = self._prep_cache
        if c is None or c.key != key:
            c = _prepare(pconv.weight, pconv.bias, conv.weight, self.p_n, dtype, need_wt_t)
            c.key = key
            self._prep_cache = c
        elif need_wt_t and c.wt_t is None:
            c.wt_t = c.wt.t().contiguous()
        return c

    def invalidate(self):
        """Drop the cached operand forms of the parameters and the folded BatchNorm.  Needed only after an in-place update
        that bypasses the version counter (`param.data.op_()`); `optimizer.step()`, `load_state_dict`, `.to()` bump it."""
        self._prep_cache = None
        self.conv[1].__dict__.pop(_FOLD_CACHE_ATTR, None)

    def _fused_ok(self, x) -> bool:
        if not self.use_fused_inference:
            return False
        bn = self.conv[1]
        if bn.running_mean is None or self.conv[0].bias is not None:
            return False
        B, C, H, W = x.shape
        return bool(_lib.load().ldconv_fused_supported(B, C, H, W, self.num_param, int(self.stride),
                                                       self.conv[0].out_channels, _DTYPES[x.dtype]))

    def forward(self, x):
        # Reduced precision at the boundary (the reference trainer runs the forward under torch.cuda.amp.autocast,
        # engine/trainer.py:693,800; validator / get_FPS.py call model.half()):
        #   bf16 autocast + fp32 input (the image, layer 0)  -> cast to bf16: the layer takes the bf16 tensor-core kernels
        #   fp16 input, or fp16 autocast + fp32 input          -> exact up-cast to fp32, fp32 kernels, fp16 result (what the
        #                                                        reference's own convs return under fp16 autocast)
        out_dtype = None
        if x.is_cuda:
            ac = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else None
            if x.dtype == torch.float32 and ac == torch.bfloat16:
                x = x.to(torch.bfloat16)
            elif x.dtype in _COMPUTE_AS or (x.dtype == torch.float32 and ac in _COMPUTE_AS):
                out_dtype = x.dtype if x.dtype in _COMPUTE_AS else ac
                x = x.to(_COMPUTE_AS[out_dtype])
        y = self._forward(x)
        return y if out_dtype is None else y.to(out_dtype)

    def _forward(self, x):
        _check_input(x)
        conv, bn = self.conv[0], self.conv[1]
        training = self.training
        if x.shape[0] == 0 and not (training and bn.track_running_stats):
            # empty batch: the reference returns an empty (0, O, h, w) tensor in eval mode (nothing to launch)
            s_ = int(self.stride)
            return x.new_empty((0, conv.out_channels, (x.shape[2] - 1) // s_ + 1, (x.shape[3] - 1) // s_ + 1))
        momentum = bn.momentum
        if training and bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
            if momentum is None:  # cumulative moving average, as torch.nn.modules.batchnorm._BatchNorm.forward
                momentum = 1.0 / float(bn.num_batches_tracked)
        if momentum is None:
            momentum = 0.0
        grad_mode = torch.is_grad_enabled() and (
            x.requires_grad or any(p.requires_grad for p in self.parameters(recurse=True)))
        prepared = self._prepared(x.dtype, grad_mode)
        if not training and not grad_mode and bn.running_mean is not None and conv.bias is None:
            return infer_nhwc(self, _nhwc(x)).permute(0, 3, 1, 2)      # lean inference path (cached BN fold)
        return _LDConvFunction.apply(x, self.p_conv.weight, self.p_conv.bias, conv.weight, conv.bias, bn.weight, bn.bias,
                                     bn.running_mean, bn.running_var, self.p_n, int(self.stride), bn.eps, momentum,
                                     training, prepared)


_FOLD_CACHE_ATTR = "_ldc_fold_cache"


def infer_nhwc(mod: "LDConv", x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Inference forward on a dense NHWC tensor (B,H,W,C) -> NHWC (B,h,w,O); `out` may be a channel slice of a wider NHWC
    buffer (the concat buffer of the consumer), which the GEMM epilogue then writes directly.  One-kernel path for the
    shapes `_fused_ok` selects, else offset conv (tensor cores when eligible) -> TMA-tiled gather -> tcgen05 GEMM."""
    L = _lib.load()
    dt = _DTYPES[x.dtype]
    conv, bn = mod.conv[0], mod.conv[1]
    B, H, W, C = x.shape
    N, s, O = mod.num_param, int(mod.stride), conv.out_channels
    h, w = (H - 1) // s + 1, (W - 1) // s + 1
    st = _stream()
    pr = mod._prepared(x.dtype, False)
    scale, shift = _folded_bn(bn, x.device)
    dense = out is None or (out.stride(2) == O and out.is_contiguous())
    if dense and mod._fused_ok(x.permute(0, 3, 1, 2)):
        if out is None:
            out = torch.empty((B, h, w, O), device=x.device, dtype=x.dtype)
        _lib.check(L.ldconv_fused_fwd(_ptr(x), _ptr(pr.w_off), _ptr(pr.b_off), _ptr(pr.pn), _ptr(pr.wt), _ptr(scale),
                                      _ptr(shift), _ptr(out), None, B, C, H, W, N, s, O, _lib.ACT_SILU, dt, st),
                   "ldconv_fused_fwd")
        return out
    M, K = B * h * w, N * C
    if out is None:
        out = torch.empty((B, h, w, O), device=x.device, dtype=x.dtype)
    ldo = out.stride(2)
    w_conv = pr.w_off_tc if s == 1 else pr.w_off_s2d
    if mod.use_onepass and w_conv is not None and L.ldconv_onepass_supported(B, C, H, W, N, s, O, ldo, dt):
        # one kernel, x read once: offsets never reach HBM
        _lib.check(L.ldconv_onepass_fwd(_ptr(x), _ptr(w_conv), _ptr(pr.b_off), _ptr(pr.pn), _ptr(pr.wt), _ptr(scale), _ptr(shift),
                                        _ptr(out), ldo, None, B, C, H, W, N, s, O, _lib.ACT_SILU, dt, st), "ldconv_onepass_fwd")
        return out
    off = offset_conv_nhwc(x, pr, N, s)
    if mod.use_gather_gemm and L.ldconv_gather_gemm_supported(B, C, H, W, N, s, O, ldo, dt):
        # gather + GEMM + BN + SiLU in one persistent kernel: the (M, N*C) operand never reaches HBM
        _lib.check(L.ldconv_gather_gemm_fwd(_ptr(x), _ptr(off), _ptr(pr.pn), _ptr(pr.wt), _ptr(scale), _ptr(shift), _ptr(out), ldo,
                                            B, C, H, W, N, s, O, _lib.ACT_SILU, dt, st), "ldconv_gather_gemm_fwd")
        return out
    operand = torch.empty((M, K), device=x.device, dtype=x.dtype)
    _lib.check(L.ldconv_gather_fwd(_ptr(x), _ptr(off), _ptr(pr.pn), _ptr(operand), None, None, B, C, H, W, N, s, dt, st),
               "ldconv_gather_fwd")
    if dt == _lib.BF16 and K % 8 == 0 and O <= 256:
        _lib.check(L.ldconv_conv1x1_bn_act_fwd(_ptr(operand), K, _ptr(pr.wt), _ptr(scale), _ptr(shift), None, 0, _ptr(out), ldo,
                                               M, K, O, _lib.ACT_SILU, dt, st), "ldconv_conv1x1_bn_act_fwd")
    else:
        tmp = out if ldo == O else torch.empty((B, h, w, O), device=x.device, dtype=x.dtype)
        _lib.check(L.ldconv_gemm_fwd(_ptr(operand), _ptr(pr.wt), _ptr(scale), _ptr(shift), _ptr(tmp), None, None, None, M, K, O,
                                     _lib.ACT_SILU, dt, st), "ldconv_gemm_fwd")
        if tmp is not out:
            out.copy_(tmp)
    return out


def _folded_bn(bn: nn.BatchNorm2d, device):
    """Eval-mode BatchNorm folded to per-channel scale/shift through the C ABI (eps read from the module).  Cached on the
    BatchNorm module while its parameters / statistics / eps are unchanged (the fold is five tiny launches otherwise)."""
    key = (bn.eps, _ver(bn.running_mean), _ver(bn.running_var), _ver(bn.weight), _ver(bn.bias), str(device),
           torch.is_inference_mode_enabled())
    cached = bn.__dict__.get(_FOLD_CACHE_ATTR)
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    scale, shift = _fold_bn_uncached(bn, device)
    bn.__dict__[_FOLD_CACHE_ATTR] = (key, scale, shift)
    return scale, shift


def _fold_bn_uncached(bn: nn.BatchNorm2d, device):
    L = _lib.load()
    O = bn.num_features
    scale = torch.empty(O, device=device, dtype=torch.float32)
    shift = torch.empty(O, device=device, dtype=torch.float32)
    gamma = None if bn.weight is None else bn.weight.detach().float().contiguous()
    beta = None if bn.bias is None else bn.bias.detach().float().contiguous()
    rm, rv = bn.running_mean.float(), bn.running_var.float()
    _lib.check(L.ldconv_bn_finalize(None, None, 0, _ptr(gamma), _ptr(beta), _ptr(rm), _ptr(rv), float(bn.eps), 0.0, 0,
                                    _ptr(scale), _ptr(shift), None, None, O, _stream()), "ldconv_bn_finalize")
    return scale, shift


# ---- training-mode BatchNorm2d + SiLU of a `Conv` block through the library's kernels -------------------------------------------
class _BNSiLUTrainFunction(torch.autograd.Function):
    """act(BatchNorm2d(pre)) in TRAINING mode for a bf16 channels_last pre-activation (the `Conv` block of the reference,
    nn/modules/conv.py:41-59, after its convolution): batch statistics (fp64 sums) -> scale / shift + running statistics ->
    one apply pass; backward = one reduction pass + one apply pass.  Four passes over the tensor instead of ATen's batch_norm
    (collect, transform, backward reduce, backward elementwise) + SiLU forward / backward (28 % of the config-4 step)."""

    @staticmethod
    def forward(ctx, pre, gamma, beta, running_mean, running_var, eps, momentum):
        L = _lib.load()
        st = _stream()
        B, O, H, W = pre.shape
        M = B * H * W
        ph = _nhwc(pre)                                   # (B,H,W,O) dense: zero-copy for channels_last
        dev = pre.device
        stats = torch.zeros((2, O), device=dev, dtype=torch.float64)
        _lib.check(L.ldconv_col_stats(_ptr(ph), _ptr(stats[0]), _ptr(stats[1]), M, O, _lib.BF16, st), "ldconv_col_stats")
        scale = torch.empty(O, device=dev, dtype=torch.float32)
        shift = torch.empty(O, device=dev, dtype=torch.float32)
        mean = torch.empty(O, device=dev, dtype=torch.float32)
        invstd = torch.empty(O, device=dev, dtype=torch.float32)
        g32, b32 = gamma.detach().float().contiguous(), beta.detach().float().contiguous()
        _lib.check(L.ldconv_bn_finalize(_ptr(stats[0]), _ptr(stats[1]), M, _ptr(g32), _ptr(b32), _ptr(running_mean), _ptr(running_var),
                                        float(eps), float(momentum), 1, _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd), O, st),
                   "ldconv_bn_finalize")
        out = torch.empty((B, H, W, O), device=dev, dtype=pre.dtype)
        _lib.check(L.ldconv_bn_act_apply(_ptr(ph), _ptr(scale), _ptr(shift), _ptr(out), M, O, _lib.ACT_SILU, _lib.BF16, st),
                   "ldconv_bn_act_apply")
        ctx.save_for_backward(ph, scale, shift, mean, invstd)
        ctx.gdtype = gamma.dtype
        return out.permute(0, 3, 1, 2)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        L = _lib.load()
        st = _stream()
        ph, scale, shift, mean, invstd = ctx.saved_tensors
        B, H, W, O = ph.shape
        M = B * H * W
        go = _nhwc(grad_out.to(ph.dtype))
        red = torch.zeros((2, O), device=ph.device, dtype=torch.float64)
        _lib.check(L.ldconv_bn_act_bwd_reduce(_ptr(ph), _ptr(go), _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd), _ptr(red), M, O,
                                              _lib.ACT_SILU, _lib.BF16, st), "ldconv_bn_act_bwd_reduce")
        gpre = torch.empty_like(ph)
        _lib.check(L.ldconv_bn_act_bwd_apply(_ptr(ph), _ptr(go), _ptr(scale), _ptr(shift), _ptr(mean), _ptr(invstd), _ptr(red), _ptr(gpre),
                                             M, O, _lib.ACT_SILU, 1, _lib.BF16, st), "ldconv_bn_act_bwd_apply")
        return gpre.permute(0, 3, 1, 2), red[1].to(ctx.gdtype), red[0].to(ctx.gdtype), None, None, None, None


def bn_silu_train(pre: torch.Tensor, bn: nn.BatchNorm2d) -> torch.Tensor:
    """Training-mode `SiLU(bn(pre))` through the library (see _BNSiLUTrainFunction), or None when the case is not covered (the
    caller then runs the torch modules): needs a CUDA bf16 tensor, affine BatchNorm2d with fp32 running statistics, > 1 value
    per channel."""
    if not (pre.is_cuda and pre.dtype == torch.bfloat16 and pre.dim() == 4 and bn.training and bn.affine and bn.track_running_stats
            and bn.running_mean is not None and bn.running_mean.dtype == torch.float32 and pre.numel() > pre.shape[1]):
        return None
    momentum = bn.momentum
    if bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
        if momentum is None:
            momentum = 1.0 / float(bn.num_batches_tracked)
    return _BNSiLUTrainFunction.apply(pre, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, momentum)


# ---- the reference's plugin hook -------------------------------------------------------------------------------------------
_REF_MODULES = ("ultralytics.nn.modules.conv", "ultralytics.nn.modules", "ultralytics.nn.modules.block",
                "ultralytics.nn.tasks")


def install(verbose: bool = False):
    """Rebind the name `LDConv` in the four reference modules that hold it (conv.py:28, modules/__init__.py:62,
    modules/block.py:9, nn/tasks.py:11) so that `parse_model` (nn/tasks.py:813, `globals()[m]`) instantiates this class
    for every `LDConv` YAML row and `Bottleneck_LDConv` / `C2f_LDConv` (block.py:611-677) pick it up too.
    Returns the list of modules patched; a no-op (empty list) when ultralytics is not importable."""
    import importlib
    patched = []
    for name in _REF_MODULES:
        try:
            mod = importlib.import_module(name)
        except Exception:  # ultralytics absent (the GPU box): nothing to patch
            continue
        if hasattr(mod, "LDConv"):
            setattr(mod, "LDConv", LDConv)
            patched.append(name)
    if verbose:
        print(f"experiment_yolo_b200.install(): LDConv rebound in {patched}")
    return patched


def convert(model: nn.Module) -> nn.Module:
    """In-place conversion of an already-built reference model: every module whose class is named `LDConv` and has the
    reference's children gets this class (parameters, buffers and state_dict untouched)."""
    for m in model.modules():
        if type(m).__name__ == "LDConv" and not isinstance(m, LDConv) and hasattr(m, "p_conv") and hasattr(m, "p_n"):
            m.__class__ = LDConv
            m._prep_cache = None
            # the reference's no-op backward hook wraps p_conv in BackwardHookFunction nodes; drop it
            m.p_conv._backward_hooks.clear()
    return model
