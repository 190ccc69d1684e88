"""Fused bf16 inference executor for the DEAL-YOLO-LD graph (SURVEY.md 8f "next" rows, widened after the LDConv path).

`FusedDealYolo(model)` takes an eval-mode `dealyolo.DealYolo` -- or the reference's own `DetectionModel` built from
yolov8-LD-P2.yaml (rows are recognised by class name and structure) -- on a B200 and runs the same graph with the library's kernels
instead of the eager torch modules around LDConv:

  * every `Conv` block (Conv2d no-bias + BatchNorm2d + SiLU, reference nn/modules/conv.py:41-59) is ONE kernel: the
    BatchNorm is folded to a per-channel scale/shift (what the reference's own `model.fuse()` does for these blocks,
    nn/tasks.py:168-195) and applied with the SiLU in the epilogue of the tcgen05 GEMM (1x1) or of the tcgen05 implicit-GEMM
    3x3 kernel;
  * C2f / SPPF (nn/modules/block.py:151-232) write their branches straight into the concatenated NHWC buffer through the
    kernels' pixel-stride arguments, so the `chunk` / `torch.cat` copies disappear; Bottleneck's residual add rides in the
    3x3 kernel's epilogue;
  * ScalSeq (nn/extra_modules/block.py:3414-3443): Conv3d(1x1x1)+BatchNorm3d+LeakyReLU(0.1) is point-wise per level, so it is
    evaluated at each level's NATIVE resolution (nearest up-sampling commutes with point-wise ops) as a folded 1x1 GEMM with a
    LeakyReLU epilogue; only the final max over the three levels runs at P2 resolution;
  * Detect (nn/modules/head.py:43-93): the 3x3 / 1x1 convs through the same kernels, and the DFL softmax-expectation +
    dist2bbox + stride scaling + class sigmoid of all three levels in one decode kernel;
  * LDConv layers call the module's own inference path (fused tcgen05 / small-C kernel, or offset-conv -> gather -> GEMM).

Activations are dense NHWC bf16 tensors (or channel slices of them).  No CPU / eager fallback for the fused ops: a missing
library or an unsupported device raises.  Numerics: bf16 storage, fp32 accumulation; tests/test_gpu_model.py compares with
the golden output of the reference DetectionModel.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .ldconv import LDConv, convert, infer_nhwc

_ACT = {"none": _lib.ACT_NONE, "silu": _lib.ACT_SILU, "leaky": _lib.ACT_LEAKY01}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _nhwc_geometry(t: torch.Tensor):
    """(B,H,W,C) tensor whose channels are contiguous and whose pixels are `ld` elements apart (a dense NHWC tensor or a
    channel slice of one)."""
    B, H, W, C = t.shape
    ld = t.stride(2)
    assert t.stride(3) == 1 and t.stride(1) == W * ld and t.stride(0) == H * W * ld, "not an NHWC (slice) tensor"
    return B, H, W, C, ld


def _is_conv_block(m) -> bool:
    """the reference's `Conv` block (nn/modules/conv.py:41-59): Conv2d + BatchNorm2d + SiLU, or Conv2d(bias) + SiLU once
    `model.fuse()` (nn/tasks.py:168-195) has folded the BatchNorm away"""
    return (isinstance(getattr(m, "conv", None), nn.Conv2d) and isinstance(getattr(m, "act", None), nn.SiLU)
            and (not hasattr(m, "bn") or isinstance(m.bn, nn.BatchNorm2d)))


class _Folded:
    """Conv weight in the kernels' (Cout, K) bf16 layout + folded per-channel affine (fp32)."""
    __slots__ = ("w", "scale", "shift", "k", "stride", "cin", "cout", "_packed_cache")

    def __init__(self, conv: nn.Conv2d, bn=None, extra_bias=None):
        w = conv.weight.detach()
        self.cout, self.cin = w.shape[0], w.shape[1]
        self.k = w.shape[2] if w.dim() >= 4 else 1
        self.stride = conv.stride[0] if isinstance(conv.stride, tuple) else conv.stride
        if w.dim() == 5:                       # Conv3d(1,1,1)
            w = w.reshape(self.cout, self.cin, 1, 1)
            self.k = 1
        if self.k == 1:
            self.w = w.reshape(self.cout, self.cin).to(torch.bfloat16).contiguous()
        else:
            self.w = w.permute(0, 2, 3, 1).reshape(self.cout, self.k * self.k * self.cin).to(torch.bfloat16).contiguous()
        dev = w.device
        bias = conv.bias.detach().float() if conv.bias is not None else torch.zeros(self.cout, device=dev)
        if bn is not None:
            inv = torch.rsqrt(bn.running_var.detach().float() + bn.eps)
            g = bn.weight.detach().float() if bn.weight is not None else torch.ones_like(inv)
            b = bn.bias.detach().float() if bn.bias is not None else torch.zeros_like(inv)
            self.scale = (g * inv).contiguous()
            self.shift = (b + (bias - bn.running_mean.detach().float()) * g * inv).contiguous()
        else:
            self.scale = torch.ones(self.cout, device=dev)
            self.shift = bias.contiguous()


def _fold(block) -> _Folded:
    return _Folded(block.conv, getattr(block, "bn", None))


def _pack_factor(C, O, rows, ldx, ldo):
    """pixels per GEMM row for a narrow dense 1x1 conv (ldconv_conv1x1_bn_act_packed_fwd), 1 = the plain kernel"""
    if ldx != C or C > _GEMM_PACK_MAX_C or O < 16 or (O & (O - 1)) or C % 8 or ldo % 8:
        return 1
    for P in (4, 2):
        if P * C <= _GEMM_PACK_MAX_K and P * O <= 128 and rows % P == 0:
            return P
    return 1


def _packed(p: _Folded, P: int):
    """block-diagonal weights and repeated affine of a folded 1x1 conv, cached on the _Folded"""
    cache = getattr(p, "_packed_cache", None)
    if cache is None:
        cache = p._packed_cache = {}
    if P not in cache:
        cache[P] = (torch.block_diag(*([p.w.float()] * P)).to(torch.bfloat16).contiguous(), p.scale.repeat(P).contiguous(),
                    p.shift.repeat(P).contiguous())
    return cache[P]


def conv1x1(x: torch.Tensor, p: _Folded, out: torch.Tensor, act: str = "silu", residual=None, out2=None, c2_lo=0):
    """out2: optional dense (B,H,W,c) tensor that also receives the output channels [c2_lo, c2_lo + c) (same pass)."""
    B, H, W, C, ldx = _nhwc_geometry(x)
    _, _, _, O, ldo = _nhwc_geometry(out)
    assert C == p.cin and O == p.cout
    if residual is None and out2 is None:
        P = _pack_factor(C, O, B * H * W, ldx, ldo)
        if P > 1:
            wp, sc, sh = _packed(p, P)
            _lib.check(_lib.load().ldconv_conv1x1_bn_act_packed_fwd(
                x.data_ptr(), wp.data_ptr(), sc.data_ptr(), sh.data_ptr(), out.data_ptr(), ldo, B * H * W, C, O, P, _ACT[act],
                _lib.BF16, _stream()), "ldconv_conv1x1_bn_act_packed_fwd")
            return out
    ldr = _nhwc_geometry(residual)[4] if residual is not None else 0
    if out2 is not None:
        _, _, _, c2, ld2 = _nhwc_geometry(out2)
        _lib.check(_lib.load().ldconv_conv1x1_bn_act_fwd2(
            x.data_ptr(), ldx, p.w.data_ptr(), p.scale.data_ptr(), p.shift.data_ptr(),
            None if residual is None else residual.data_ptr(), ldr, out.data_ptr(), ldo, out2.data_ptr(), ld2, c2_lo, c2, B * H * W,
            C, O, _ACT[act], _lib.BF16, _stream()), "ldconv_conv1x1_bn_act_fwd2")
        return out
    _lib.check(_lib.load().ldconv_conv1x1_bn_act_fwd(
        x.data_ptr(), ldx, p.w.data_ptr(), p.scale.data_ptr(), p.shift.data_ptr(),
        None if residual is None else residual.data_ptr(), ldr, out.data_ptr(), ldo, B * H * W, C, O, _ACT[act], _lib.BF16,
        _stream()), "ldconv_conv1x1_bn_act_fwd")
    return out


def conv3x3(x: torch.Tensor, p: _Folded, out: torch.Tensor, act: str = "silu", residual=None):
    B, H, W, C, ldx = _nhwc_geometry(x)
    _, h, w, O, ldo = _nhwc_geometry(out)
    assert C == p.cin and O == p.cout and h == (H - 1) // p.stride + 1 and w == (W - 1) // p.stride + 1
    ldr = _nhwc_geometry(residual)[4] if residual is not None else 0
    _lib.check(_lib.load().ldconv_conv3x3_bn_act_fwd(
        x.data_ptr(), ldx, p.w.data_ptr(), p.scale.data_ptr(), p.shift.data_ptr(),
        None if residual is None else residual.data_ptr(), ldr, out.data_ptr(), ldo, B, C, H, W, O, p.stride, _ACT[act],
        _lib.BF16, _stream()), "ldconv_conv3x3_bn_act_fwd")
    return out


def _new(like: torch.Tensor, B, H, W, C):
    return torch.empty((B, H, W, C), device=like.device, dtype=torch.bfloat16)


# Settled by measurement in round 1 (the A/B environment switches are gone):
#   * C2f.cv1 does NOT also store the first Bottleneck's chunk densely (0.7 % slower in the step: the slice is L2-resident);
#   * a C2f whose output also feeds a later Concat writes that slice from its last conv (second output of the GEMM);
#   * the SSFF maximum + Add run in the epilogue of the finest level's point-wise GEMM;
#   * an LDConv whose other consumers are up-samplings writes straight into the Concat buffer (they read the slice);
#   * the first convs of the two Detect branches of a level run as one stacked conv;
#   * narrow dense 1x1 convs (Cin <= 48) run with P pixels per GEMM row and block-diagonal weights: -19...-25 % stand-alone at
#     Cin = 32 / 48, nothing at 64 where the plain kernel's warp-staged stores gain 10 % (profiles/r1_gemm_pack_ab_s4.jsonl).
_GEMM_PACK_MAX_K = 256
_GEMM_PACK_MAX_C = 48


class _C2f:
    def __init__(self, m):
        self.c = m.c
        self.n = len(m.m)
        self.cv1 = _fold(m.cv1)
        self.cv2 = _fold(m.cv2)
        self.blocks = [(_fold(b.cv1), _fold(b.cv2), b.add) for b in m.m]

    def __call__(self, x, out2=None):
        """out2: optional NHWC slice (of a later Concat's buffer) that receives a second copy of the output in the same pass"""
        B, H, W, _ = x.shape
        c, n = self.c, self.n
        cat = _new(x, B, H, W, (2 + n) * c)
        conv1x1(x, self.cv1, cat[..., : 2 * c])
        tmp = _new(x, B, H, W, c)
        for i, (p1, p2, add) in enumerate(self.blocks):
            src = cat[..., (1 + i) * c: (2 + i) * c]
            conv3x3(src, p1, tmp)
            conv3x3(tmp, p2, cat[..., (2 + i) * c: (3 + i) * c], residual=src if add else None)
        return conv1x1(cat, self.cv2, _new(x, B, H, W, self.cv2.cout), out2=out2, c2_lo=0)


class _SPPF:
    def __init__(self, m):
        self.cv1 = _fold(m.cv1)
        self.cv2 = _fold(m.cv2)
        self.k = m.m.kernel_size

    def __call__(self, x):
        B, H, W, _ = x.shape
        c = self.cv1.cout
        cat = _new(x, B, H, W, 4 * c)
        conv1x1(x, self.cv1, cat[..., :c])
        _lib.check(_lib.load().ldconv_sppf_pools(cat.data_ptr(), cat[..., c:].data_ptr(), cat[..., 2 * c:].data_ptr(),
                                                 cat[..., 3 * c:].data_ptr(), 4 * c, B, H, W, c, self.k, _lib.BF16, _stream()),
                   "ldconv_sppf_pools")
        return conv1x1(cat, self.cv2, _new(x, B, H, W, self.cv2.cout))


def upsample_into(x: torch.Tensor, out: torch.Tensor, factor: int):
    """nearest up-sampling of an NHWC tensor (slice) straight into an NHWC tensor (slice)."""
    B, H, W, C, ldx = _nhwc_geometry(x)
    ldo = _nhwc_geometry(out)[4]
    _lib.check(_lib.load().ldconv_upsample_nearest(x.data_ptr(), ldx, out.data_ptr(), ldo, B, H, W, C, factor, _lib.BF16,
                                                   _stream()), "ldconv_upsample_nearest")
    return out


def add_nhwc(xs):
    """`Add` row: the sum of ALL its inputs (NHWC tensors / channel slices), fp32 accumulation, one rounding per launch
    (up to four inputs per launch; longer lists chain through the running sum)."""
    import ctypes
    pending = list(xs)
    B, H, W, C, _ = _nhwc_geometry(pending[0])
    while True:
        part, pending = pending[:4], pending[4:]
        srcs = (ctypes.c_void_p * len(part))(*[t.data_ptr() for t in part])
        lds = (ctypes.c_int * len(part))(*[_nhwc_geometry(t)[4] for t in part])
        out = _new(part[0], B, H, W, C)
        _lib.check(_lib.load().ldconv_add_nhwc(srcs, lds, len(part), out.data_ptr(), C, B * H * W, C, _lib.BF16, _stream()),
                   "ldconv_add_nhwc")
        if not pending:
            return out
        pending.insert(0, out)


class _ScalSeq:
    def __init__(self, m):
        self.conv0 = _fold(m.conv0) if hasattr(m, "conv0") else None
        self.conv1 = _fold(m.conv1)
        self.conv2 = _fold(m.conv2)
        self.mix = _Folded(m.conv3d, m.bn)          # Conv3d(1x1x1) + BatchNorm3d, point-wise per level

    def __call__(self, xs, addend=None):
        fine, mid, coarse = xs
        ch = self.mix.cout
        if self.conv0 is not None:
            fine = conv1x1(fine, self.conv0, _new(fine, *fine.shape[:3], ch))
        mid = conv1x1(mid, self.conv1, _new(mid, *mid.shape[:3], ch))
        coarse = conv1x1(coarse, self.conv2, _new(coarse, *coarse.shape[:3], ch))
        B, H, W, _ = fine.shape
        out = _new(fine, B, H, W, ch)
        if ch % 16 == 0:
            # the finest level's point-wise GEMM takes the maximum over the levels (and the Add) in its epilogue: its own
            # (B,H,W,ch) map (105 MB at P2, batch 64) never reaches HBM and the tail kernel disappears
            z1 = conv1x1(mid, self.mix, _new(mid, *mid.shape[:3], ch), act="leaky")
            z2 = conv1x1(coarse, self.mix, _new(coarse, *coarse.shape[:3], ch), act="leaky")
            _, _, _, C, ldx = _nhwc_geometry(fine)
            p = self.mix
            _lib.check(_lib.load().ldconv_conv1x1_bn_act_maxup_fwd(
                fine.data_ptr(), ldx, p.w.data_ptr(), p.scale.data_ptr(), p.shift.data_ptr(), z1.data_ptr(), z1.shape[1], z1.shape[2],
                z2.data_ptr(), z2.shape[1], z2.shape[2], None if addend is None else addend.data_ptr(),
                0 if addend is None else _nhwc_geometry(addend)[4], out.data_ptr(), ch, B, H, W, C, ch, _ACT["leaky"], _lib.BF16,
                _stream()), "ldconv_conv1x1_bn_act_maxup_fwd")
            return out
        z = [conv1x1(t, self.mix, _new(t, *t.shape[:3], ch), act="leaky") for t in (fine, mid, coarse)]
        # MaxPool3d((3,1,1)) over the stacked depth axis (+ the following Add layer when it consumes this output)
        _lib.check(_lib.load().ldconv_scalseq_tail(
            z[0].data_ptr(), z[1].data_ptr(), z[2].data_ptr(), None if addend is None else addend.data_ptr(),
            0 if addend is None else _nhwc_geometry(addend)[4], out.data_ptr(), ch, B, H, W, z[1].shape[1], z[1].shape[2],
            z[2].shape[1], z[2].shape[2], ch, _lib.BF16, _stream()), "ldconv_scalseq_tail")
        return out


def conv1x1_detect(x: torch.Tensor, p: _Folded, y: torch.Tensor, mode: int, nc: int, stride: float, a0: int):
    """last 1x1 conv of a Detect branch with the decode in its epilogue (ldconv_conv1x1_detect_fwd): mode 1 = box branch -> rows
    0..3 of y (B, 4+nc, anchors), mode 2 = class branch -> rows 4..4+nc; this level's anchors start at column a0"""
    B, H, W, C, ldx = _nhwc_geometry(x)
    assert C == p.cin and y.is_contiguous() and y.shape[1] == 4 + nc
    _lib.check(_lib.load().ldconv_conv1x1_detect_fwd(x.data_ptr(), ldx, p.w.data_ptr(), p.scale.data_ptr(), p.shift.data_ptr(),
                                                      y.data_ptr(), mode, B, H, W, C, p.cout, nc, float(stride), a0, y.shape[2],
                                                      _lib.BF16, _stream()), "ldconv_conv1x1_detect_fwd")


class _Detect:
    # The six conv chains of the head (3 pyramid levels x {box, class} branch) are independent until the decode: they are forked
    # onto side streams (captured as parallel branches of the CUDA graph) so that the small P3 / P4 kernels and the two
    # branches of a level share the GPU instead of running back to back.  Every buffer is allocated on the calling stream
    # before the fork, the side streams only launch kernels, and the join precedes any reuse.
    parallel_branches = True
    # The last 1x1 conv of every branch decodes its own accumulator rows (DFL expectation + dist2bbox + stride / sigmoid) straight
    # into y: no logits in HBM, no decode launches; bit-identical to the unfused sequence.  The raw head maps (`feats`, what the
    # reference returns beside y in eval mode) are then not produced: set False when they are needed.
    decode_in_epilogue = True

    def __init__(self, m):
        self.nc, self.reg_max, self.no = m.nc, m.reg_max, m.no
        self.stride = [float(s) for s in m.stride]
        self.box = [[_fold(s[0]), _fold(s[1]), _Folded(s[2])] for s in m.cv2]
        self.cls = [[_fold(s[0]), _fold(s[1]), _Folded(s[2])] for s in m.cv3]
        self.streams = None
        # The first convs of the two branches of a level read the same input (nn/modules/head.py:43-52: cv2[i][0] ch -> 64, cv3[i][0]
        # ch -> 32).  As ONE conv with the weights stacked (Cout = 96) the input is read once and the MMAs run at N = 96: the wide
        # convs are tensor-pipe-bound and a 128 x N x 16 MMA costs the same pipe time for any N <= 128.  The second
        # convs then read channel slices of the 96-channel buffer.  Levels with more than 64 input channels keep two convs (the
        # stacked weights no longer fit the zero-copy kernel's shared memory).
        self.first = []
        for b, c in zip(self.box, self.cls):
            f = None
            if b[0].cin <= 64 and (b[0].cout + c[0].cout) % 16 == 0 and b[0].cout % 16 == 0:
                f = object.__new__(_Folded)
                f.w = torch.cat([b[0].w, c[0].w], 0).contiguous()
                f.scale = torch.cat([b[0].scale, c[0].scale]).contiguous()
                f.shift = torch.cat([b[0].shift, c[0].shift]).contiguous()
                f.k, f.stride, f.cin, f.cout = b[0].k, b[0].stride, b[0].cin, b[0].cout + c[0].cout
            self.first.append(f)

    @staticmethod
    def _chain(x, ps, bufs, skip_first=False, det=None):
        t = x if skip_first else conv3x3(x, ps[0], bufs[0])
        t = conv3x3(t, ps[1], bufs[1])
        if det is not None:
            return conv1x1_detect(t, ps[2], *det)
        return conv1x1(t, ps[2], bufs[2], act="none")

    def __call__(self, xs):
        L = _lib.load()
        B = xs[0].shape[0]
        total = sum(x.shape[1] * x.shape[2] for x in xs)
        y = torch.empty((B, 4 + self.nc, total), device=xs[0].device, dtype=torch.bfloat16)
        fuse_dec = self.decode_in_epilogue and self.reg_max == 16 and self.nc <= 16 and all(b[2].cin % 8 == 0 for b in self.box + self.cls)
        a0s, acc = [], 0
        for x in xs:
            a0s.append(acc)
            acc += x.shape[1] * x.shape[2]
        dets = [((y, 1, self.nc, self.stride[lvl], a0s[lvl]), (y, 2, self.nc, self.stride[lvl], a0s[lvl])) if fuse_dec else (None, None)
                for lvl in range(len(xs))]
        bufs = []
        for lvl, x in enumerate(xs):
            _, H, W, _ = x.shape
            if self.first[lvl] is not None:      # one buffer for the stacked first conv; the branches read its channel slices
                both = _new(x, B, H, W, self.first[lvl].cout)
                b0, c0 = both[..., : self.box[lvl][0].cout], both[..., self.box[lvl][0].cout:]
            else:
                both, b0, c0 = None, _new(x, B, H, W, self.box[lvl][0].cout), _new(x, B, H, W, self.cls[lvl][0].cout)
            bb = [b0, _new(x, B, H, W, self.box[lvl][1].cout), None if fuse_dec else _new(x, B, H, W, 4 * self.reg_max)]
            cb = [c0, _new(x, B, H, W, self.cls[lvl][1].cout), None if fuse_dec else _new(x, B, H, W, self.nc)]
            bufs.append((bb, cb, both))
        cur = torch.cuda.current_stream(xs[0].device)
        if self.parallel_branches:
            if self.streams is None:
                self.streams = [torch.cuda.Stream(xs[0].device) for _ in range(2 * len(xs))]
            stacked_done = {}
            for k, st in enumerate(self.streams):
                lvl, br = k // 2, k % 2
                fused = self.first[lvl] is not None
                if br == 0:
                    st.wait_stream(cur)
                    with torch.cuda.stream(st):
                        if fused:
                            conv3x3(xs[lvl], self.first[lvl], bufs[lvl][2])
                            stacked_done[lvl] = torch.cuda.Event()
                            stacked_done[lvl].record(st)      # the class branch starts here, not after the whole box chain
                        self._chain(bufs[lvl][0][0] if fused else xs[lvl], self.box[lvl], bufs[lvl][0], skip_first=fused, det=dets[lvl][0])
                else:
                    if fused:
                        st.wait_event(stacked_done[lvl])
                    else:
                        st.wait_stream(cur)
                    with torch.cuda.stream(st):
                        self._chain(bufs[lvl][1][0] if fused else xs[lvl], self.cls[lvl], bufs[lvl][1], skip_first=fused, det=dets[lvl][1])
            for st in self.streams:
                cur.wait_stream(st)
        else:
            for lvl, x in enumerate(xs):
                fused = self.first[lvl] is not None
                if fused:
                    conv3x3(x, self.first[lvl], bufs[lvl][2])
                self._chain(bufs[lvl][0][0] if fused else x, self.box[lvl], bufs[lvl][0], skip_first=fused, det=dets[lvl][0])
                self._chain(bufs[lvl][1][0] if fused else x, self.cls[lvl], bufs[lvl][1], skip_first=fused, det=dets[lvl][1])
        if fuse_dec:
            return y, None
        feats, a0 = [], 0
        for lvl, x in enumerate(xs):
            _, H, W, _ = x.shape
            b, c = bufs[lvl][0][2], bufs[lvl][1][2]
            _lib.check(L.ldconv_detect_decode(b.data_ptr(), c.data_ptr(), y.data_ptr(), B, H, W, self.nc, self.reg_max,
                                              self.stride[lvl], a0, total, _lib.BF16, _stream()), "ldconv_detect_decode")
            feats.append((b, c))
            a0 += H * W
        return y, feats


class _Deferred:
    """an nn.Upsample output that has not been materialised yet"""
    __slots__ = ("src", "factor")

    def __init__(self, src, factor):
        self.src, self.factor = src, factor


class FusedDealYolo:
    """Inference executor: `y, feats = FusedDealYolo(model)(images)` with `images` (B,3,H,W) bf16 (any memory format).
    `y` is the decoded (B, 4+nc, anchors) tensor of the reference's Detect head in eval mode."""

    def __init__(self, model):
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("FusedDealYolo needs the model on a CUDA device (sm_100a); there is no CPU path")
        _lib.check(_lib.load().ldconv_device_check(), "ldconv_device_check")
        if model.training:
            raise RuntimeError("FusedDealYolo is an inference executor: call model.eval() first")
        convert(model)      # reference LDConv rows -> this package's class, in place (parameters / state_dict untouched)
        self.model = model
        self.layers = []
        builders = {"c2f": _C2f, "sppf": _SPPF, "scalseq": _ScalSeq, "detect": _Detect}
        for layer, kind in zip(model.model, self.recognise(model)):
            if kind in builders:
                op = ("fn", builders[kind](layer))
            elif kind == "ldconv":
                op = ("ldconv", layer)
            elif kind == "up":
                op = ("up", int(layer.scale_factor))
            else:
                op = (kind, None)
            self.layers.append((op, layer.f, layer.i))
        self.save = set(model.save)
        # LDConv whose ONLY consumer is a Concat writes its output straight into that Concat's buffer
        self.cat_plan = {}
        self.dual_plan = {}
        consumers = {}
        for (kind, arg), f, i in self.layers:
            for j in ([f] if isinstance(f, int) else f):
                consumers.setdefault((i + j) if j < 0 else j, []).append(i)
        chan_out = {}
        for (kind, arg), f, i in self.layers:
            if kind == "ldconv":
                chan_out[i] = arg.conv[0].out_channels
        for (kind, arg), f, i in self.layers:
            if kind != "cat":
                continue
            srcs = [(i + j) if j < 0 else j for j in f]
            widths = []
            for sidx in srcs:
                k2, a2 = self.layers[sidx][0]
                if k2 == "ldconv":
                    widths.append(a2.conv[0].out_channels)
                elif k2 == "up":
                    up_src = self.layers[sidx][1]
                    up_src = (sidx + up_src) if up_src < 0 else up_src
                    widths.append(self._out_channels(up_src))
                else:
                    widths.append(self._out_channels(sidx))
            if any(wd is None for wd in widths):
                continue
            c0 = 0
            for sidx, wd in zip(srcs, widths):
                k2, a2 = self.layers[sidx][0]
                # other consumers may be up-samplings: they read the slice through its pixel stride (yaml row 8 -> 9 and 22)
                others_up = all(self.layers[cj][0][0] == "up" for cj in consumers.get(sidx, []) if cj != i)
                if k2 == "ldconv" and c0 % 8 == 0 and sidx not in self.cat_plan and (
                        consumers.get(sidx) == [i] or others_up):
                    self.cat_plan[sidx] = (i, c0, sum(widths))
                elif (k2 == "fn" and isinstance(a2, _C2f) and c0 % 8 == 0 and wd % 16 == 0
                      and sidx not in self.dual_plan):
                    # a C2f with further consumers (yaml row 12 -> 13 and 19): its last 1x1 conv writes the dense output AND the
                    # Concat's slice in one pass instead of a strided copy_ later (62 us for 52 MB at P3)
                    self.dual_plan[sidx] = (i, c0, sum(widths))
                c0 += wd
        # ScalSeq -> Add (yolov8-LD-P2.yaml rows 24, 25): the Add's other operand is folded into the ScalSeq tail kernel
        for n, ((kind, arg), f, i) in enumerate(self.layers[:-1]):
            (k2, _), f2, _ = self.layers[n + 1]
            if kind == "fn" and isinstance(arg, _ScalSeq) and k2 == "add" and isinstance(f2, list) and -1 in f2 and len(f2) == 2:
                other = [j for j in f2 if j != -1][0]
                self.layers[n] = (("scalseq_add", (arg, other)), f, i)
                self.layers[n + 1] = (("identity", None), -1, self.layers[n + 1][2])

    @staticmethod
    def recognise(model):
        """Row kinds of a yolov8-LD-P2-style graph.  Rows are recognised by class NAME and structure, not by type: the
        reference's own `DetectionModel` (nn/tasks.py:275; modules from nn/modules/{conv,block,head}.py and
        nn/extra_modules/block.py) has the same attribute names as this package's dealyolo graph, so either can be handed
        in.  Needs no device (tests/test_host_cpu.py runs it on the real reference model)."""
        kinds = []
        for layer in model.model:
            name = type(layer).__name__
            if name == "LDConv" and hasattr(layer, "p_conv") and hasattr(layer, "p_n"):
                kinds.append("ldconv")
            elif name == "C2f" and _is_conv_block(layer.cv1) and all(type(b).__name__ == "Bottleneck" for b in layer.m):
                kinds.append("c2f")
            elif name == "SPPF" and _is_conv_block(layer.cv1):
                kinds.append("sppf")
            elif name == "ScalSeq" and hasattr(layer, "conv3d"):
                kinds.append("scalseq")
            elif name == "Detect" and hasattr(layer, "cv2") and hasattr(layer, "cv3") and hasattr(layer, "dfl"):
                kinds.append("detect")
            elif name == "Concat" and getattr(layer, "d", 1) == 1:
                kinds.append("cat")
            elif name == "Add":
                kinds.append("add")
            elif isinstance(layer, nn.Upsample) and layer.mode == "nearest" and float(layer.scale_factor).is_integer():
                kinds.append("up")
            else:
                raise NotImplementedError(f"FusedDealYolo: no fused executor for {name}")
        return kinds

    def _out_channels(self, idx):
        kind, arg = self.layers[idx][0]
        if kind == "ldconv":
            return arg.conv[0].out_channels
        if kind == "fn" and hasattr(arg, "cv2"):
            return arg.cv2.cout
        if kind in ("fn", "scalseq_add") :
            a = arg[0] if kind == "scalseq_add" else arg
            return a.mix.cout if hasattr(a, "mix") else None
        if kind == "up":
            f = self.layers[idx][1]
            return self._out_channels((idx + f) if f < 0 else f)
        return None

    @staticmethod
    def _concat(xs, buf=None):
        """channel concat of NHWC tensors; deferred up-samplings are written straight into their slice, inputs that already
        live in `buf` (producers that wrote into the concat buffer) are skipped"""
        ref = next(t for t in xs if isinstance(t, torch.Tensor))
        B = ref.shape[0]
        H, W = (ref.shape[1], ref.shape[2])
        chans = [t.src.shape[3] if isinstance(t, _Deferred) else t.shape[3] for t in xs]
        out = buf if buf is not None else _new(ref, B, H, W, sum(chans))
        c0 = 0
        for t, c in zip(xs, chans):
            dst = out[..., c0:c0 + c]
            if isinstance(t, _Deferred):
                upsample_into(t.src, dst, t.factor)
            elif t.data_ptr() != dst.data_ptr():
                upsample_into(t, dst, 1)          # factor 1 = a strided slice copy
            c0 += c
        return out

    # Images per pass through the graph.  The batch is cut into micro-batches so that the producer -> consumer tensors of
    # neighbouring kernels (105 MB per 32-channel P2 map at 64 images) stay resident in the 126 MB L2 instead of making a round
    # trip through HBM; None = the whole batch at once.  Outputs are written into one preallocated result.
    micro_batch = None

    @torch.no_grad()
    def __call__(self, images: torch.Tensor):
        mb = self.micro_batch
        if mb and images.shape[0] > mb:
            ys, feats = [], []
            for b0 in range(0, images.shape[0], mb):
                y, f = self._forward(images[b0:b0 + mb])
                ys.append(y)
                feats.append(f)
            return torch.cat(ys, 0), feats
        return self._forward(images)

    def _forward(self, images: torch.Tensor):
        if images.dtype == torch.uint8:       # raw (B,C,H,W) uint8 batch: normalise + NHWC in one kernel
            B, C, H, W = images.shape
            x = torch.empty((B, H, W, C), device=images.device, dtype=torch.bfloat16)
            _lib.check(_lib.load().ldconv_image_u8_to_nhwc(images.contiguous().data_ptr(), x.data_ptr(), B, C, H, W, 1.0 / 255.0,
                                                           _lib.BF16, _stream()), "ldconv_image_u8_to_nhwc")
        else:
            x = images.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous()   # NHWC; zero-copy for channels_last input
        saved = []
        cat_bufs = {}
        for (kind, arg), f, i in self.layers:
            if f != -1:
                x = saved[f] if isinstance(f, int) else [x if j == -1 else saved[j] for j in f]
            if kind == "ldconv":
                plan = self.cat_plan.get(i)
                if plan is None:
                    x = infer_nhwc(arg, x)
                else:       # the only consumer is a Concat: write straight into its buffer
                    cat_i, c0, ctot = plan
                    B, H, W, _ = x.shape
                    s_ = int(arg.stride)
                    h, w = (H - 1) // s_ + 1, (W - 1) // s_ + 1
                    buf = cat_bufs.get(cat_i)
                    if buf is None:
                        buf = cat_bufs[cat_i] = _new(x, B, h, w, ctot)
                    O = arg.conv[0].out_channels
                    x = infer_nhwc(arg, x, out=buf[..., c0:c0 + O])
            elif kind == "scalseq_add":     # ScalSeq whose only consumer is the next Add layer: one tail kernel does both
                x = arg[0](x, addend=saved[arg[1]])
            elif kind == "fn":
                plan = self.dual_plan.get(i)
                if plan is None:
                    x = arg(x)
                else:
                    cat_i, c0, ctot = plan
                    B, H, W, _ = x.shape
                    buf = cat_bufs.get(cat_i)
                    if buf is None:
                        buf = cat_bufs[cat_i] = _new(x, B, H, W, ctot)
                    x = arg(x, out2=buf[..., c0:c0 + arg.cv2.cout])
            elif kind == "cat":
                buf = cat_bufs.pop(i, None)
                for pos, j in enumerate(f):          # inputs a C2f already wrote into the buffer: hand _concat the slice itself
                    plan = self.dual_plan.get((i + j) if j < 0 else j)
                    if plan is not None and plan[0] == i and buf is not None:
                        x[pos] = buf[..., plan[1]:plan[1] + x[pos].shape[3]]
                x = self._concat(x, buf)
            elif kind == "add":
                x = add_nhwc(x)
            elif kind == "up":
                x = _Deferred(x, arg)        # materialised by the consumer (Concat writes it straight into its buffer)
            elif kind == "identity":
                pass
            saved.append(x if i in self.save else None)
        return x


class PipelinedPredictor:
    """The end-to-end inference call on HOST buffers: `submit(u8_batch_in_pinned_memory)` / `result()`.

    Mirrors what the reference's predictor does per batch (upload uint8 images, normalise on the device, forward,
    non_max_suppression, read detections back; engine/predictor.py:120-140,266-300, utils/ops.py:292) with the B200-side
    plumbing around the fused executor: the whole forward (uint8 -> bf16 NHWC conversion and, with `nms=`, the device-side
    NMS included) is captured once per input slot in a CUDA graph, and the slots are rotated so that the upload of batch i+1
    and the download of batch i-1 overlap the compute of batch i on separate streams.  Results come back in submission order.

    nms=None: `result()` is the decoded head output (B, 4+nc, anchors) bf16 (43 MB per 64 images at 640x640).
    nms=dict(conf_thres=..., iou_thres=..., agnostic=..., max_det=...): `result()` is (detections (B, max_det, 6) fp32 rows
    (x1, y1, x2, y2, conf, cls), counts (B,) int32) -- 0.46 MB per 64 images; `detections()` slices them per image."""

    def __init__(self, model, batch: int, imgsz: int, channels: int = 3, slots: int = 2, nms: Optional[dict] = None):
        from . import nms as _nms
        self.exec = FusedDealYolo(model)
        dev = next(model.parameters()).device
        self.dev, self.slots = dev, slots
        self.nms = dict(nms) if nms is not None else None
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.u8 = [torch.zeros((batch, channels, imgsz, imgsz), device=dev, dtype=torch.uint8) for _ in range(slots)]
        self.graphs, self.y_dev, self.y_host = [], [], []
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                y, _ = self.exec(self.u8[0])
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        # The NMS launch is NOT part of the captured forward: it runs on the download stream, so the 64 one-CTA-per-image
        # sequential loops of batch i overlap the forward of batch i+1 instead of holding 84 SMs idle at the tail of every graph
        # (inside the graph the end-to-end rate was 5 % below the raw-output variant)
        self._nms_fn = _nms.nms_padded
        self.nms_ws = None
        if self.nms is not None:
            self.nms_ws = _nms.nms_workspace(batch, y.shape[2], int(self.nms.get("max_nms", 30000)), dev)      # launches are ordered on s_out
        self.y_raw = []
        for k in range(slots):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                y, _ = self.exec(self.u8[k])
            self.graphs.append(g)
            self.y_raw.append(y)
            if self.nms is not None:
                md = int(self.nms.get("max_det", 300))
                ys = (torch.zeros((batch, md, 6), device=dev, dtype=torch.float32), torch.zeros((batch,), device=dev, dtype=torch.int32))
            else:
                ys = (y,)
            self.y_dev.append(ys)
            self.y_host.append(tuple(torch.empty(tuple(t.shape), dtype=t.dtype).pin_memory() for t in ys))
        self.ev_in = [torch.cuda.Event() for _ in range(slots)]
        self.ev_comp = [torch.cuda.Event() for _ in range(slots)]
        self.ev_out = [torch.cuda.Event() for _ in range(slots)]
        self.n_submitted = 0
        self.n_returned = 0
        self.h2d_bytes = self.u8[0].numel()
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.y_host[0])

    def submit(self, host_u8: torch.Tensor):
        """host_u8: (B,C,H,W) uint8 in pinned host memory.  Asynchronous."""
        k = self.n_submitted % self.slots
        comp = torch.cuda.current_stream(self.dev)
        if self.n_submitted >= self.slots:
            self.s_in.wait_event(self.ev_comp[k])        # slot's previous batch has been consumed by its graph
        else:
            self.s_in.wait_stream(comp)
        with torch.cuda.stream(self.s_in):
            self.u8[k].copy_(host_u8, non_blocking=True)
            self.ev_in[k].record(self.s_in)
        comp.wait_event(self.ev_in[k])
        if self.n_submitted >= self.slots:
            comp.wait_event(self.ev_out[k])              # slot's previous result has left the device buffer
        self.graphs[k].replay()
        self.ev_comp[k].record(comp)
        self.s_out.wait_event(self.ev_comp[k])
        with torch.cuda.stream(self.s_out):
            if self.nms is not None:
                self._nms_fn(self.y_raw[k], out=self.y_dev[k][0], count=self.y_dev[k][1], workspace=self.nms_ws, **self.nms)
            for h, d in zip(self.y_host[k], self.y_dev[k]):
                h.copy_(d, non_blocking=True)
            self.ev_out[k].record(self.s_out)
        self.n_submitted += 1

    def result(self):
        """Blocks until the oldest outstanding batch's result is in host memory and returns it (pinned buffers that are
        reused `slots` submissions later): the decoded head output, or (detections, counts) with `nms=`."""
        assert self.n_returned < self.n_submitted, "no outstanding batch"
        k = self.n_returned % self.slots
        self.ev_out[k].synchronize()
        self.n_returned += 1
        return self.y_host[k][0] if self.nms is None else self.y_host[k]

    @staticmethod
    def detections(result):
        """(detections, counts) of `result()` -> list of (n_i, 6) tensors, the return value of the reference's
        non_max_suppression (utils/ops.py:427)"""
        det, cnt = result
        cnt = cnt.tolist()
        if any(c < 0 for c in cnt):
            raise RuntimeError("nms: more candidates than max_nms in an image; raise conf_thres")
        return [det[b, :c] for b, c in enumerate(cnt)]

    def drain_to(self, stream=None):
        """make `stream` (default: current) wait for every outstanding download (for device-side timing)"""
        stream = stream or torch.cuda.current_stream(self.dev)
        for k in range(min(self.slots, self.n_submitted)):
            stream.wait_event(self.ev_out[k])
