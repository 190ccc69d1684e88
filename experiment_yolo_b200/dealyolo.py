"""The DEAL-YOLO-LD graph around LDConv: what bench.py's headline metric (640x640 images/s) runs.

This is the benchmark harness, not a re-implementation of Ultralytics: it builds the model of
/root/reference/ultralytics/cfg/models/yolov8-LD-P2.yaml from a YAML in the reference's own row format
(`[from, repeats, module, args]`, cfg/deal-yolo-ld-p2.yaml) with the reference's name-lookup hook
(nn/tasks.py:813 `globals()[m]` -> MODULES[name] here), and names every sub-module like the reference so that a reference
`state_dict()` loads with strict=True.  The neighbours of LDConv (SURVEY.md 8f "next" rows) are plain torch modules
here (cuDNN): Conv / C2f / Bottleneck / SPPF (nn/modules/conv.py:41-59, block.py:151-171,209-232,320-335),
ScalSeq / Add (nn/extra_modules/block.py:3414-3443,3479-3484), Detect / DFL (nn/modules/head.py:19-93,
block.py:37-56).  `ldconv_cls` selects the LDConv implementation: the CUDA module (default) or, for the CPU baseline
leg only, the eager port under oracle/ (injected by bench.py / tests, never imported from here).
"""
from __future__ import annotations

import math
import os
from copy import deepcopy

import torch
import torch.nn as nn
import torch.nn.functional as F
import yaml

from .ldconv import LDConv

DEFAULT_CFG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg", "deal-yolo-ld-p2.yaml")


def _same_pad(k):
    return k // 2 if isinstance(k, int) else [v // 2 for v in k]


class Conv(nn.Module):
    """Conv2d(no bias) + BatchNorm2d + SiLU (reference nn/modules/conv.py:41-59)."""

    def __init__(self, c1, c2, k=1, s=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, _same_pad(k), bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = nn.SiLU() if act is True else (act if isinstance(act, nn.Module) else nn.Identity())

    # training on the GPU under bf16 autocast: batch-stat BatchNorm + SiLU run as four passes of the library's kernels
    # (ldconv.bn_silu_train) instead of ATen's batch_norm + SiLU forward / backward; class switch for A/B and tests
    fused_bn_silu_train = True

    def forward(self, x):
        y = self.conv(x)
        if self.training and self.fused_bn_silu_train and y.is_cuda and y.dtype == torch.bfloat16 and isinstance(self.act, nn.SiLU) \
                and torch.is_grad_enabled():
            from .ldconv import bn_silu_train
            z = bn_silu_train(y, self.bn)
            if z is not None:
                return z
        return self.act(self.bn(y))


class Bottleneck(nn.Module):
    """Two 3x3 Convs with an optional residual (reference nn/modules/block.py:320-335, as used by C2f: e=1.0)."""

    def __init__(self, c1, c2, shortcut=True, k=(3, 3), e=0.5):
        super().__init__()
        hidden = int(c2 * e)
        self.cv1 = Conv(c1, hidden, k[0], 1)
        self.cv2 = Conv(hidden, c2, k[1], 1)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    """CSP block with n bottlenecks whose outputs are all concatenated (reference nn/modules/block.py:209-232)."""

    def __init__(self, c1, c2, n=1, shortcut=False, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut, k=(3, 3), e=1.0) for _ in range(n))

    def forward(self, x):
        parts = list(self.cv1(x).chunk(2, 1))
        for blk in self.m:
            parts.append(blk(parts[-1]))
        return self.cv2(torch.cat(parts, 1))


class SPPF(nn.Module):
    """Three chained 5x5 max-pools concatenated with their input (reference nn/modules/block.py:151-171)."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        self.cv1 = Conv(c1, c1 // 2, 1, 1)
        self.cv2 = Conv(c1 // 2 * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)

    def forward(self, x):
        x = self.cv1(x)
        p1 = self.m(x)
        p2 = self.m(p1)
        return self.cv2(torch.cat((x, p1, p2, self.m(p2)), 1))


class Upsample(nn.Upsample):
    """nn.Upsample whose nearest / integer-factor case on a CUDA bf16 tensor runs through the library (train_ops.upsample_nearest:
    bf16 in, bf16 out, one kernel each way); everything else is torch's op.  Same constructor and state (none) as nn.Upsample,
    so the YAML row `nn.Upsample [None, 2, nearest]` (yolov8-LD-P2.yaml:26,33) builds it unchanged."""

    def forward(self, x):
        if self.mode == "nearest" and self.scale_factor is not None and float(self.scale_factor).is_integer():
            if torch.is_autocast_enabled("cuda") and x.is_cuda and x.dtype == torch.float32 and torch.get_autocast_dtype("cuda") == torch.bfloat16:
                x = x.to(torch.bfloat16)
            from .train_ops import upsample_nearest
            y = upsample_nearest(x, int(self.scale_factor))
            if y is not None:
                return y
        return super().forward(x)


def _nearest_to(x, size):
    """F.interpolate(x, size, mode="nearest") -- through the library when `size` is an integer multiple of x's size"""
    H, W = x.shape[2:]
    if size[0] % H == 0 and size[1] % W == 0 and size[0] // H == size[1] // W:
        from .train_ops import upsample_nearest
        y = upsample_nearest(x, size[0] // H)
        if y is not None:
            return y
    return F.interpolate(x, size, mode="nearest")


class Concat(nn.Module):
    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, xs):
        return torch.cat(xs, self.d)


class Add(nn.Module):
    """Element-wise sum of the inputs (reference nn/extra_modules/block.py:3479-3484)."""

    def forward(self, xs):
        if xs[0].is_cuda and xs[0].dim() == 4:
            from .train_ops import add_maps
            y = add_maps(xs)
            if y is not None:
                return y
        return torch.sum(torch.stack(xs, dim=0), dim=0)


class ScalSeq(nn.Module):
    """SSFF: 1x1 Convs bring three pyramid levels to `channel`, the coarser two are nearest-upsampled to the finest, the
    three maps are stacked along a depth axis, Conv3d(1x1x1)+BatchNorm3d+LeakyReLU(0.1), then max over the depth axis
    (reference nn/extra_modules/block.py:3414-3443)."""

    # training on the GPU under bf16 autocast: the tail after the three 1x1 Convs through the library (train_ops.scalseq_tail);
    # class switch for A/B and tests
    fused_train_tail = True

    def __init__(self, inc, channel):
        super().__init__()
        if channel != inc[0]:
            self.conv0 = Conv(inc[0], channel, 1)
        self.conv1 = Conv(inc[1], channel, 1)
        self.conv2 = Conv(inc[2], channel, 1)
        self.conv3d = nn.Conv3d(channel, channel, kernel_size=(1, 1, 1))
        self.bn = nn.BatchNorm3d(channel)
        self.act = nn.LeakyReLU(0.1)
        self.pool_3d = nn.MaxPool3d(kernel_size=(3, 1, 1))

    def forward(self, xs):
        fine, mid, coarse = xs
        if hasattr(self, "conv0"):
            fine = self.conv0(fine)
        size = fine.shape[2:]
        mid, coarse = self.conv1(mid), self.conv2(coarse)
        if self.training and self.fused_train_tail and fine.is_cuda and mid.dtype == torch.bfloat16 and torch.is_grad_enabled() \
                and isinstance(self.act, nn.LeakyReLU) and abs(self.act.negative_slope - 0.1) < 1e-12:
            from .train_ops import scalseq_tail
            # an fp32 finest map under bf16 autocast (no conv0 in front of it): the Conv3d would round it to bf16 anyway
            f16 = fine.to(torch.bfloat16) if fine.dtype == torch.float32 and torch.is_autocast_enabled("cuda") else fine
            y = scalseq_tail(f16, mid, coarse, self.conv3d, self.bn)
            if y is not None:
                return y
        mid = _nearest_to(mid, size)
        coarse = _nearest_to(coarse, size)
        vol = torch.stack([fine, mid, coarse], dim=2)                # (B, C, 3, H, W)
        vol = self.act(self.bn(self.conv3d(vol)))
        return self.pool_3d(vol).squeeze(2)


class DFL(nn.Module):
    """Expectation over the 16-bin distance distribution, as a frozen 1x1 conv (reference nn/modules/block.py:37-56)."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1

    def forward(self, x):
        b, _, a = x.shape
        return self.conv(x.view(b, 4, self.c1, a).transpose(2, 1).softmax(1)).view(b, 4, a)


def make_anchors(feats, strides, offset=0.5):
    """Cell-centre anchor points and their strides for every level (reference utils/tal.py:294-307)."""
    pts, strs = [], []
    for f, s in zip(feats, strides):
        h, w = f.shape[2:]
        sx = torch.arange(w, device=f.device, dtype=f.dtype) + offset
        sy = torch.arange(h, device=f.device, dtype=f.dtype) + offset
        gy, gx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((gx, gy), -1).view(-1, 2))
        strs.append(torch.full((h * w, 1), float(s), dtype=f.dtype, device=f.device))
    return torch.cat(pts), torch.cat(strs)


class Detect(nn.Module):
    """Three-level decoupled head with DFL decode (reference nn/modules/head.py:19-93)."""

    def __init__(self, nc=80, ch=()):
        super().__init__()
        self.nc, self.nl, self.reg_max = nc, len(ch), 16
        self.no = nc + 4 * self.reg_max
        self.stride = torch.zeros(self.nl)
        c2 = max(16, ch[0] // 4, 4 * self.reg_max)
        c3 = max(ch[0], min(nc, 100))
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(c, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for c in ch)
        self.cv3 = nn.ModuleList(nn.Sequential(Conv(c, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, nc, 1)) for c in ch)
        self.dfl = DFL(self.reg_max)
        self._shape, self.anchors, self.strides = None, None, None

    def forward(self, xs):
        xs = [torch.cat((self.cv2[i](x), self.cv3[i](x)), 1) for i, x in enumerate(xs)]
        if self.training:
            return xs
        shape = xs[0].shape
        flat = torch.cat([x.reshape(shape[0], self.no, -1) for x in xs], 2)
        if self._shape != (shape, xs[0].dtype, xs[0].device):
            a, s = make_anchors(xs, self.stride, 0.5)
            self.anchors, self.strides = a.transpose(0, 1), s.transpose(0, 1)
            self._shape = (shape, xs[0].dtype, xs[0].device)
        box, cls = flat.split((4 * self.reg_max, self.nc), 1)
        lt, rb = self.dfl(box).chunk(2, 1)
        anc = self.anchors.unsqueeze(0)
        x1y1, x2y2 = anc - lt, anc + rb
        dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * self.strides
        return torch.cat((dbox, cls.sigmoid()), 1), xs

    def bias_init(self):
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / s) ** 2)


def _divisible(x, d=8):
    return math.ceil(x / d) * d


class DealYolo(nn.Module):
    """YAML -> layer list -> forward over saved outputs, like DetectionModel (reference nn/tasks.py:275-333, 85-126)."""

    def __init__(self, cfg=DEFAULT_CFG, ch=3, nc=None, ldconv_cls=LDConv, strides=(4.0, 8.0, 16.0)):
        super().__init__()
        spec = cfg if isinstance(cfg, dict) else yaml.safe_load(open(cfg))
        spec = deepcopy(spec)
        if nc:
            spec["nc"] = nc
        self.yaml = spec
        modules = {"LDConv": ldconv_cls, "C2f": C2f, "SPPF": SPPF, "Conv": Conv, "Concat": Concat, "Add": Add,
                   "ScalSeq": ScalSeq, "Detect": Detect, "nn.Upsample": Upsample}
        depth, width, max_ch = spec["scales"][spec.get("scale") or next(iter(spec["scales"]))]
        chans, layers, save = [ch], [], set()
        for i, (f, n, name, args) in enumerate(spec["backbone"] + spec["head"]):
            m = modules[name]                                   # the YAML hook: module resolved by name
            args = [spec["nc"] if a == "nc" else (None if a == "None" else a) for a in args]
            n = max(round(n * depth), 1) if n > 1 else n
            if name in ("LDConv", "C2f", "SPPF", "Conv"):
                c1, c2 = chans[f], _divisible(min(args[0], max_ch) * width)
                args = [c1, c2, *args[1:]]
                if name == "C2f":
                    args.insert(2, n)
                    n = 1
            elif name == "Concat":
                c2 = sum(chans[x] for x in f)
            elif name == "Add":
                c2 = chans[f[-1]]
            elif name == "ScalSeq":
                c2 = _divisible(args[0] * width)
                args = [[chans[x] for x in f], c2]
            elif name == "Detect":
                args.append([chans[x] for x in f])
                c2 = None
            else:
                c2 = chans[f]
            layer = nn.Sequential(*(m(*args) for _ in range(n))) if n > 1 else m(*args)
            layer.i, layer.f = i, f
            save.update(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
            layers.append(layer)
            if i == 0:
                chans = []
            chans.append(c2)
        self.model = nn.Sequential(*layers)
        self.save = sorted(save)
        head = self.model[-1]
        # the reference discovers these with a 2x3x640x640 probe forward (nn/tasks.py:309-324); they are the P2/P3/P4
        # strides of this graph, so they are set directly and checked by tests/test_torch_port.py::test_strides_agree_with_a_probe_forward against a probe
        head.stride = torch.tensor(strides)
        self.stride = head.stride
        head.bias_init()
        for mod in self.modules():                              # utils/torch_utils.py:342-352
            if type(mod) is nn.BatchNorm2d:
                mod.eps, mod.momentum = 1e-3, 0.03
            elif isinstance(mod, (nn.SiLU, nn.LeakyReLU)):
                mod.inplace = True

    def forward(self, x):
        saved = []
        for layer in self.model:
            if layer.f != -1:
                x = saved[layer.f] if isinstance(layer.f, int) else [x if j == -1 else saved[j] for j in layer.f]
            x = layer(x)
            saved.append(x if layer.i in self.save else None)
        return x

    def ldconv_layers(self):
        return [m for m in self.model if type(m).__name__.startswith("LDConv")]


def channels_last_(model: nn.Module) -> nn.Module:
    """Put every 4-D parameter in torch.channels_last (NHWC) so cuDNN keeps activations NHWC end to end and LDConv gets
    its dense-NHWC inputs zero-copy.  (nn.Module.to(memory_format=...) also touches the 5-D Conv3d weight of ScalSeq and
    raises, so the conversion is done per parameter.)"""
    for p in model.parameters():
        if p.dim() == 4:
            p.data = p.data.contiguous(memory_format=torch.channels_last)
    return model


def seeded_state(model: nn.Module, seed: int = 0, p_conv_sigma: float = 0.05):
    """Deterministic synthetic weights that do not depend on module construction order: every floating tensor of the
    state_dict, in key order, is drawn from one CPU generator (BatchNorm statistics / scales kept positive, LDConv
    offset-conv weights N(0, p_conv_sigma) so the sampling grid is irregular -- the shipped zero-init makes offsets
    constant).  Used for golden fixtures and synthetic benchmarks on both the reference and this model."""
    g = torch.Generator().manual_seed(seed)
    sd = model.state_dict()
    out = {}
    for k, v in sd.items():
        if not v.is_floating_point() or k.endswith("dfl.conv.weight"):
            out[k] = v.clone()
        elif k.endswith("running_var"):
            out[k] = (torch.rand(v.shape, generator=g) * 0.5 + 0.75).to(v.dtype)
        elif k.endswith("running_mean"):
            out[k] = (torch.randn(v.shape, generator=g) * 0.1).to(v.dtype)
        elif k.endswith("p_conv.weight"):
            out[k] = (torch.randn(v.shape, generator=g) * p_conv_sigma).to(v.dtype)
        elif k.endswith(("bn.weight", "conv.1.weight")) and v.dim() == 1:
            out[k] = (torch.rand(v.shape, generator=g) * 0.5 + 0.75).to(v.dtype)
        elif v.dim() == 1:
            out[k] = (torch.randn(v.shape, generator=g) * 0.1).to(v.dtype)
        else:
            fan_in = v[0].numel()
            out[k] = (torch.randn(v.shape, generator=g) * (1.0 / math.sqrt(fan_in))).to(v.dtype)
    return out
