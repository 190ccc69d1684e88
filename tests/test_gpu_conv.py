"""GPU: the tcgen05 3x3 implicit-GEMM convolution -- as LDConv's offset conv (conv.py:356,368) and as the
Conv2d + folded BatchNorm + SiLU block next to it (nn/modules/conv.py:41-59) -- against fp32 references on bf16-rounded
operands (fp32 accumulation on both sides: products of bf16 values are exact in fp32)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from experiment_yolo_b200 import _lib
from oracle import oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _p(t):
    return None if t is None else t.data_ptr()


def _st():
    return torch.cuda.current_stream().cuda_stream


def _rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


@pytest.mark.parametrize("C,N,s,H,W,B", [(16, 3, 2, 40, 56, 2), (32, 3, 2, 37, 41, 2), (64, 1, 1, 20, 20, 2), (128, 1, 1, 17, 33, 1),
                                         (64, 3, 2, 24, 24, 2), (32, 5, 1, 16, 48, 1), (16, 9, 2, 33, 19, 2), (256, 9, 1, 12, 12, 1),
                                         (48, 2, 1, 9, 9, 3), (32, 1, 1, 160, 160, 2)])
def test_offset_conv_tensor_core_vs_oracle(C, N, s, H, W, B):
    L = _lib.load()
    assert L.ldconv_offset_conv_tc_supported(C, N, s, _lib.BF16) == 1
    g = torch.Generator().manual_seed(C + N)
    x = torch.randn(B, C, H, W, generator=g).bfloat16()
    w = (torch.randn(2 * N, C, 3, 3, generator=g) * 0.1).bfloat16()
    b = torch.randn(2 * N, generator=g)
    want = oracle.offset_conv(x.float().numpy(), w.float().numpy(), b.numpy(), N, s)            # (B,2N,h,w)
    h, wo = (H - 1) // s + 1, (W - 1) // s + 1
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    wd = w.permute(0, 2, 3, 1).reshape(2 * N, 9 * C).contiguous().to(DEV)
    bd = b.to(DEV)
    off = torch.full((B, h, wo, 2 * N), float("nan"), device=DEV)
    _lib.check(L.ldconv_offset_conv_tc_fwd(_p(xd), _p(wd), _p(bd), _p(off), B, C, H, W, N, s, _lib.BF16, _st()))
    torch.cuda.synchronize()
    got = off.cpu().numpy().transpose(0, 3, 1, 2)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= 2e-4 * max(1.0, float(np.abs(want).max()))


@pytest.mark.parametrize("C,N,H,W,B", [(16, 3, 40, 56, 2), (32, 3, 38, 42, 2), (16, 1, 16, 24, 1), (32, 5, 160, 160, 2), (16, 8, 2, 2, 3),
                                       (16, 3, 320, 320, 2), (64, 3, 40, 40, 2), (64, 3, 80, 80, 5), (64, 1, 18, 34, 1)])
def test_offset_conv_stride2_space_to_depth_vs_oracle(C, N, H, W, B):
    """ldconv_offset_conv_s2d_fwd: the stride-2 offset conv (conv.py:356,368) as a zero-copy tcgen05 GEMM on the space-to-depth
    view (5-D TMA map), weights scattered by the module's own _prepare -- against the CPU oracle's offset conv and against the
    im2col tensor-core kernel on the same inputs."""
    from experiment_yolo_b200.ldconv import _prepare, base_grid
    L = _lib.load()
    assert L.ldconv_offset_conv_s2d_supported(C, N, H, W, _lib.BF16) == 1
    g = torch.Generator().manual_seed(C + N + H)
    x = torch.randn(B, C, H, W, generator=g).bfloat16()
    w = (torch.randn(2 * N, C, 3, 3, generator=g) * 0.1).bfloat16()
    b = torch.randn(2 * N, generator=g).bfloat16().float()      # _prepare rounds the bias through the activation dtype
    want = oracle.offset_conv(x.float().numpy(), w.float().numpy(), b.numpy(), N, 2)            # (B,2N,h,w)
    h, wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    pr = _prepare(w.to(DEV), b.to(DEV), torch.zeros(8, C, N, 1, device=DEV), base_grid(N), torch.bfloat16, False)
    assert pr.w_off_s2d is not None
    xd = x.permute(0, 2, 3, 1).contiguous().to(DEV)
    off = torch.full((B, h, wo, 2 * N), float("nan"), device=DEV)
    _lib.check(L.ldconv_offset_conv_s2d_fwd(_p(xd), _p(pr.w_off_s2d), _p(pr.b_off), _p(off), B, C, H, W, N, _lib.BF16, _st()))
    torch.cuda.synchronize()
    got = off.cpu().numpy().transpose(0, 3, 1, 2)
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= 2e-4 * max(1.0, float(np.abs(want).max()))
    off2 = torch.empty_like(off)
    _lib.check(L.ldconv_offset_conv_tc_fwd(_p(xd), _p(pr.w_off_tc), _p(pr.b_off), _p(off2), B, C, H, W, N, 2, _lib.BF16, _st()))
    torch.cuda.synchronize()
    assert float((off - off2).abs().max()) <= 2e-4 * max(1.0, float(off2.abs().max()))


@pytest.mark.parametrize("Cin,Cout,s,H,W,B,res", [(16, 16, 1, 40, 40, 2, True), (32, 32, 1, 37, 21, 2, True), (64, 64, 1, 20, 20, 2, False),
                                                  (32, 64, 1, 24, 40, 1, False), (128, 64, 1, 17, 17, 2, False), (64, 32, 2, 40, 40, 1, False),
                                                  (16, 48, 1, 8, 8, 3, False), (64, 64, 1, 80, 80, 2, True), (256, 128, 1, 10, 12, 1, False),
                                                  (128, 64, 1, 40, 40, 64, False)])      # one input buffer, ~6 tiles per CTA (Detect P4 at batch 64)
def test_conv3x3_bn_silu_vs_torch(Cin, Cout, s, H, W, B, res):
    L = _lib.load()
    assert L.ldconv_conv3x3_supported(Cin, Cout, s, _lib.BF16) == 1
    g = torch.Generator(device=DEV).manual_seed(Cin * 3 + Cout)
    x = torch.randn((B, Cin, H, W), device=DEV, generator=g).bfloat16()
    w = (torch.randn((Cout, Cin, 3, 3), device=DEV, generator=g) * (1.0 / (3 * Cin ** 0.5))).bfloat16()
    scale = torch.rand(Cout, device=DEV, generator=g) + 0.5
    shift = torch.randn(Cout, device=DEV, generator=g) * 0.2
    torch.backends.cudnn.allow_tf32 = False
    z = F.conv2d(x.float(), w.float(), None, s, 1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    ref = F.silu(z)
    h, wo = ref.shape[2:]
    r = torch.randn((B, h, wo, Cout), device=DEV, generator=g).bfloat16() if res else None
    if res:
        ref = ref + r.permute(0, 3, 1, 2).float()
    xd = x.permute(0, 2, 3, 1).contiguous()
    wd = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    out = torch.full((B, h, wo, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    _lib.check(L.ldconv_conv3x3_bn_act_fwd(_p(xd), Cin, _p(wd), _p(scale), _p(shift), _p(r), Cout, _p(out), Cout, B, Cin, H, W,
                                           Cout, s, _lib.ACT_SILU, _lib.BF16, _st()))
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2).float()
    assert bool(torch.isfinite(got).all())
    rel = float((got - ref).norm() / ref.norm())
    assert rel <= 6e-3, rel


def test_conv3x3_channel_slices_in_and_out():
    """Input and output as channel slices of wider NHWC buffers (what makes C2f concat-free)."""
    L = _lib.load()
    B, H, W, Cin, Cout, ldx, ldo = 2, 24, 24, 32, 32, 96, 128
    g = torch.Generator(device=DEV).manual_seed(3)
    xbuf = torch.randn((B, H, W, ldx), device=DEV, generator=g).bfloat16()
    obuf = torch.zeros((B, H, W, ldo), device=DEV, dtype=torch.bfloat16)
    w = (torch.randn((Cout, Cin, 3, 3), device=DEV, generator=g) * 0.06).bfloat16()
    xs = xbuf[..., 32:64]
    ref = F.silu(F.conv2d(xs.permute(0, 3, 1, 2).float(), w.float(), None, 1, 1))
    wd = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    xptr = xbuf.data_ptr() + 32 * 2
    optr = obuf.data_ptr() + 64 * 2
    _lib.check(L.ldconv_conv3x3_bn_act_fwd(xptr, ldx, _p(wd), None, None, None, 0, optr, ldo, B, Cin, H, W, Cout, 1,
                                           _lib.ACT_SILU, _lib.BF16, _st()))
    torch.cuda.synchronize()
    got = obuf[..., 64:96].permute(0, 3, 1, 2).float()
    assert float((got - ref).norm() / ref.norm()) <= 6e-3
    assert float(obuf[..., :64].abs().max()) == 0.0 and float(obuf[..., 96:].abs().max()) == 0.0


@pytest.mark.parametrize("Cin,Cout,lo,n,rows,ldo", [(32, 32, 16, 16, 5000, 48), (64, 64, 32, 32, 777, 96), (128, 128, 64, 64, 300, 192),
                                                    (32, 64, 0, 16, 129, 64), (64, 32, 16, 16, 40000, 48)])
def test_conv1x1_second_dense_output(Cin, Cout, lo, n, rows, ldo):
    """ldconv_conv1x1_bn_act_fwd2: the 1x1 Conv block (conv.py:41-59) writing a channel slice of a wider buffer AND, in the same
    pass, a dense copy of the output channels [lo, lo + n) (C2f's second chunk, block.py:222-226): both must equal the plain call."""
    L = _lib.load()
    g = torch.Generator(device=DEV).manual_seed(Cin + Cout + rows)
    x = torch.randn((rows, Cin), device=DEV, generator=g).bfloat16()
    wt = (torch.randn((Cout, Cin), device=DEV, generator=g) * 0.2).bfloat16()
    scale = torch.rand(Cout, device=DEV, generator=g) + 0.5
    shift = torch.randn(Cout, device=DEV, generator=g) * 0.1
    ref = torch.empty((rows, Cout), device=DEV, dtype=torch.bfloat16)
    _lib.check(L.ldconv_conv1x1_bn_act_fwd(_p(x), Cin, _p(wt), _p(scale), _p(shift), None, 0, _p(ref), Cout, rows, Cin, Cout,
                                           _lib.ACT_SILU, _lib.BF16, _st()), "conv1x1")
    buf = torch.full((rows, ldo), 3.0, device=DEV, dtype=torch.bfloat16)
    out2 = torch.full((rows, n), 5.0, device=DEV, dtype=torch.bfloat16)
    _lib.check(L.ldconv_conv1x1_bn_act_fwd2(_p(x), Cin, _p(wt), _p(scale), _p(shift), None, 0, _p(buf), ldo, _p(out2), n, lo, n, rows,
                                            Cin, Cout, _lib.ACT_SILU, _lib.BF16, _st()), "conv1x1 (two outputs)")
    torch.cuda.synchronize()
    assert torch.equal(buf[:, :Cout], ref)
    assert torch.equal(out2, ref[:, lo:lo + n])
    if ldo > Cout:
        assert float((buf[:, Cout:].float() - 3.0).abs().max()) == 0.0      # the neighbouring channels are untouched
    z = (x.double() @ wt.double().t()) * scale.double() + shift.double()
    want = (z * torch.sigmoid(z)).float()
    assert float((ref.float() - want).norm() / want.norm()) <= 6e-3


@pytest.mark.parametrize("B,H,W,C,f,ldx,ldo,c0", [(2, 5, 7, 32, 2, 32, 64, 32), (1, 40, 40, 64, 2, 128, 128, 64), (3, 3, 4, 16, 3, 16, 16, 0),
                                                 (2, 1, 1, 8, 2, 24, 8, 0)])
def test_upsample_nearest_slices_bit_exact(B, H, W, C, f, ldx, ldo, c0):
    """nn.Upsample(scale_factor=f, mode='nearest') (yaml rows 9, 14) from an NHWC channel slice into an NHWC channel slice:
    a pure copy, so bit-exact; the bytes around the destination slice stay untouched."""
    L = _lib.load()
    g = torch.Generator().manual_seed(B * 100 + C)
    src = torch.randn(B, H, W, ldx, generator=g).bfloat16().to(DEV)
    dst = torch.full((B, H * f, W * f, ldo), 7.0, dtype=torch.bfloat16, device=DEV)
    x = src[..., ldx - C:]
    out = dst[..., c0:c0 + C]
    _lib.check(L.ldconv_upsample_nearest(x.data_ptr(), ldx, out.data_ptr(), ldo, B, H, W, C, f, _lib.BF16, _st()))
    want = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=f, mode="nearest").permute(0, 2, 3, 1).bfloat16()
    assert torch.equal(out, want)
    rest = torch.ones(ldo, dtype=torch.bool)
    rest[c0:c0 + C] = False
    assert bool((dst[..., rest.to(DEV)] == 7.0).all())


@pytest.mark.parametrize("B,H,W,Cin,Cout,with_add", [(2, 16, 24, 32, 32, True), (1, 40, 40, 32, 32, False), (3, 8, 12, 64, 48, True),
                                                   (2, 20, 20, 16, 16, True)])
def test_conv1x1_maxup_epilogue_bit_exact_vs_unfused(B, H, W, Cin, Cout, with_add):
    """SSFF tail (ScalSeq + Add, nn/extra_modules/block.py:3414-3443,3479-3484) fused into the finest level's point-wise GEMM
    (ldconv_conv1x1_bn_act_maxup_fwd) against the unfused sequence conv1x1 -> ldconv_scalseq_tail on the same inputs: the
    fused epilogue rounds the level's own output to bf16 before the maximum like the stored tensor, so the results are equal
    bit for bit; and against torch (nearest up-sampling, maximum, add) within bf16 rounding."""
    L = _lib.load()
    g = torch.Generator().manual_seed(B * 7 + Cout)
    x = torch.randn(B, H, W, Cin, generator=g).bfloat16().to(DEV)
    wt = (torch.randn(Cout, Cin, generator=g) * 0.2).bfloat16().to(DEV)
    scale = (torch.rand(Cout, generator=g) + 0.5).to(DEV)
    shift = (torch.randn(Cout, generator=g) * 0.1).to(DEV)
    H1, W1, H2, W2 = H // 2, W // 2, H // 4, W // 4
    z1 = torch.randn(B, H1, W1, Cout, generator=g).bfloat16().to(DEV)
    z2 = torch.randn(B, H2, W2, Cout, generator=g).bfloat16().to(DEV)
    add = torch.randn(B, H, W, Cout, generator=g).bfloat16().to(DEV) if with_add else None
    z0 = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=DEV)
    want = torch.empty_like(z0)
    got = torch.full_like(z0, float("nan"))
    _lib.check(L.ldconv_conv1x1_bn_act_fwd(_p(x), Cin, _p(wt), _p(scale), _p(shift), None, 0, _p(z0), Cout, B * H * W, Cin, Cout,
                                           _lib.ACT_LEAKY01, _lib.BF16, _st()))
    _lib.check(L.ldconv_scalseq_tail(_p(z0), _p(z1), _p(z2), _p(add), Cout if with_add else 0, _p(want), Cout, B, H, W, H1, W1, H2, W2,
                                     Cout, _lib.BF16, _st()))
    _lib.check(L.ldconv_conv1x1_bn_act_maxup_fwd(_p(x), Cin, _p(wt), _p(scale), _p(shift), _p(z1), H1, W1, _p(z2), H2, W2, _p(add),
                                                 Cout if with_add else 0, _p(got), Cout, B, H, W, Cin, Cout, _lib.ACT_LEAKY01,
                                                 _lib.BF16, _st()))
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    pre = (x.float().reshape(-1, Cin) @ wt.float().t()) * scale + shift
    lvl = F.leaky_relu(pre, 0.1).bfloat16().float().reshape(B, H, W, Cout)
    up = lambda t: F.interpolate(t.permute(0, 3, 1, 2).float(), size=(H, W), mode="nearest").permute(0, 2, 3, 1)
    ref = torch.maximum(torch.maximum(lvl, up(z1)), up(z2))
    if with_add:
        ref = ref + add.float()
    assert (got.float() - ref).abs().max().item() <= 0.02 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,H,W,nc,stride", [(2, 5, 7, 6, 4.0), (1, 40, 40, 6, 8.0), (3, 3, 3, 1, 16.0), (2, 9, 4, 11, 32.0)])
def test_detect_decode_vs_torch(B, H, W, nc, stride):
    """Detect head decode of one level (nn/modules/head.py:55-77: DFL softmax expectation over 16 bins per side,
    nn/modules/block.py:37-56; dist2bbox xywh around the cell-centre anchors, utils/tal.py:309-319; times the stride; sigmoid of the
    class logits) against the same arithmetic in torch fp32 on the bf16 logits; columns outside the level's range stay untouched."""
    L = _lib.load()
    g = torch.Generator().manual_seed(H * 13 + W)
    box = (torch.randn(B, H, W, 64, generator=g) * 2).bfloat16().to(DEV)
    cls = (torch.randn(B, H, W, nc, generator=g) * 3).bfloat16().to(DEV)
    a0, total = 5, H * W + 9
    y = torch.full((B, 4 + nc, total), 3.0, dtype=torch.bfloat16, device=DEV)
    _lib.check(L.ldconv_detect_decode(_p(box), _p(cls), _p(y), B, H, W, nc, 16, stride, a0, total, _lib.BF16, _st()))
    p = box.float().reshape(B, H * W, 4, 16).softmax(-1)
    d = (p * torch.arange(16, device=DEV, dtype=torch.float32)).sum(-1)                       # (B, A, 4) l, t, r, b
    jj, ii = torch.meshgrid(torch.arange(W, device=DEV), torch.arange(H, device=DEV), indexing="xy")
    ax, ay = jj.reshape(-1).float() + 0.5, ii.reshape(-1).float() + 0.5
    x1, y1, x2, y2 = ax - d[..., 0], ay - d[..., 1], ax + d[..., 2], ay + d[..., 3]
    want = torch.stack([(x1 + x2) / 2, (y1 + y2) / 2, x2 - x1, y2 - y1], 1) * stride              # (B, 4, A)
    got = y[:, :, a0:a0 + H * W].float()
    assert (got[:, :4] - want).abs().max().item() <= 2 ** -7 * max(1.0, want.abs().max().item())
    assert (got[:, 4:] - cls.float().reshape(B, H * W, nc).permute(0, 2, 1).sigmoid()).abs().max().item() <= 2 ** -8
    outside = torch.cat([y[:, :, :a0], y[:, :, a0 + H * W:]], 2)
    assert bool((outside == 3.0).all())


@pytest.mark.parametrize("B,H,W,C,k", [(2, 40, 40, 64, 5), (1, 7, 9, 16, 5), (3, 20, 12, 8, 3), (2, 4, 4, 32, 5)])
def test_sppf_pools_bit_exact_vs_max_pool2d(B, H, W, C, k):
    """SPPF's three cascaded MaxPool2d(k, 1, k//2) (nn/modules/block.py:166-171) written into the channel slices 1..3 of the 4C-wide
    concat buffer whose slice 0 holds the input: pure maxima, so bit-exact against torch."""
    L = _lib.load()
    g = torch.Generator().manual_seed(H * 17 + C)
    cat = torch.full((B, H, W, 4 * C), 9.0, dtype=torch.bfloat16, device=DEV)
    x = torch.randn(B, H, W, C, generator=g).bfloat16().to(DEV)
    cat[..., :C] = x
    _lib.check(L.ldconv_sppf_pools(cat.data_ptr(), cat[..., C:].data_ptr(), cat[..., 2 * C:].data_ptr(), cat[..., 3 * C:].data_ptr(),
                                   4 * C, B, H, W, C, k, _lib.BF16, _st()))
    y = x.permute(0, 3, 1, 2).float()
    for lvl in range(1, 4):
        y = F.max_pool2d(y, k, 1, k // 2)
        assert torch.equal(cat[..., lvl * C:(lvl + 1) * C], y.permute(0, 2, 3, 1).bfloat16())
    assert torch.equal(cat[..., :C], x)


@pytest.mark.parametrize("Cin,Cout,P,rows,ldo,act", [(32, 32, 4, 4096, 48, "silu"), (48, 32, 2, 1000, 32, "silu"), (64, 64, 2, 2050, 64, "none"),
                                                     (16, 16, 4, 260, 16, "leaky"), (64, 32, 4, 131072 + 4, 32, "silu")])
def test_conv1x1_pixel_packed_matches_plain_kernel(Cin, Cout, P, rows, ldo, act):
    """ldconv_conv1x1_bn_act_packed_fwd (P pixels per GEMM row, block-diagonal weights) against ldconv_conv1x1_bn_act_fwd on the same
    inputs: the extra products are exact zeros and a row's K-order is unchanged, so the outputs are equal bit for bit; the channels
    next to an output slice stay untouched; and against fp64 on the bf16 operands."""
    L = _lib.load()
    A = {"none": _lib.ACT_NONE, "silu": _lib.ACT_SILU, "leaky": _lib.ACT_LEAKY01}[act]
    g = torch.Generator().manual_seed(Cin * 3 + Cout + P)
    x = torch.randn(rows, Cin, generator=g).bfloat16().to(DEV)
    wt = (torch.randn(Cout, Cin, generator=g) / Cin ** 0.5).bfloat16().to(DEV)
    scale = (torch.rand(Cout, generator=g) + 0.5).to(DEV)
    shift = (torch.randn(Cout, generator=g) * 0.2).to(DEV)
    wp = torch.block_diag(*([wt.float()] * P)).bfloat16().contiguous()
    want = torch.full((rows, ldo), 5.0, dtype=torch.bfloat16, device=DEV)
    got = torch.full((rows, ldo), 5.0, dtype=torch.bfloat16, device=DEV)
    _lib.check(L.ldconv_conv1x1_bn_act_fwd(_p(x), Cin, _p(wt), _p(scale), _p(shift), None, 0, _p(want), ldo, rows, Cin, Cout, A,
                                           _lib.BF16, _st()))
    scale_rep, shift_rep = scale.repeat(P).contiguous(), shift.repeat(P).contiguous()      # kept alive across the asynchronous call
    _lib.check(L.ldconv_conv1x1_bn_act_packed_fwd(_p(x), _p(wp), _p(scale_rep), _p(shift_rep), _p(got), ldo, rows, Cin, Cout, P, A,
                                                  _lib.BF16, _st()))
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    z = (x.double() @ wt.double().t()) * scale.double() + shift.double()
    ref = {"none": z, "silu": z * torch.sigmoid(z), "leaky": torch.where(z > 0, z, 0.1 * z)}[act]
    assert _rel(got[:, :Cout].float().cpu().numpy(), ref.cpu().numpy()) <= 6e-3
    if ldo > Cout:
        assert bool((got[:, Cout:] == 5.0).all())


@pytest.mark.parametrize("B,H,W", [(3, 64, 96), (2, 17, 23), (1, 640, 640)])
def test_image_u8_to_nhwc_bit_exact(B, H, W):
    """uint8 NCHW batch -> bf16 NHWC in [0, 1] (the reference predictor's `im.half(); im /= 255` + layout change,
    engine/predictor.py:120-131): the 8-pixels-per-thread kernel (H*W % 8 == 0) and the per-pixel fallback, bit for bit."""
    from experiment_yolo_b200 import _lib
    g = torch.Generator().manual_seed(H)
    u8 = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=g).to(DEV)
    out = torch.empty((B, H, W, 3), device=DEV, dtype=torch.bfloat16)
    _lib.check(_lib.load().ldconv_image_u8_to_nhwc(u8.data_ptr(), out.data_ptr(), B, 3, H, W, 1.0 / 255.0, _lib.BF16,
                                                   torch.cuda.current_stream().cuda_stream), "ldconv_image_u8_to_nhwc")
    want = (u8.float() * torch.tensor(1.0 / 255.0, dtype=torch.float32)).bfloat16().permute(0, 2, 3, 1).contiguous()
    assert torch.equal(out, want)
