"""Eager-PyTorch CPU port of the reference LDConv -- the CPU baseline that bench.py times.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/ldconv_oracle.c's header): never imported by the product package.

The reference implementation of the hot path IS eager PyTorch (/root/reference/ultralytics/nn/modules/conv.py:350-503)
and cannot travel to the GPU box, so the `cpu_baseline` / `--impl reference` legs time this port instead
(`kind: "port"`).  To keep the timing honest it executes the same sequence of ATen ops as the reference: CPU meshgrid for
p_0 every call (conv.py:435-444), ~a dozen element-wise passes for floor / clamp / weights (:375-393), four
`torch.gather` calls over an int64 index expanded across channels (:456-489), the weighted sum (:402-405), the
rearrange copy (:494-503) and the (N,1) Conv2d + BatchNorm2d + SiLU (:355,408).  tests/test_torch_port.py checks that
it reproduces the golden vectors bit-for-bit in the forward pass and that its state_dict layout equals the
reference's; oracle/gen_model_golden.py (authoring container only) asserts that the whole graph built on this port equals the
real reference DetectionModel bit for bit and records both wall-clock times.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class LDConvTorchPort(nn.Module):
    def __init__(self, inc, outc, num_param, stride=1, bias=None):
        super().__init__()
        self.num_param, self.stride = num_param, stride
        self.conv = nn.Sequential(
            nn.Conv2d(inc, outc, kernel_size=(num_param, 1), stride=(num_param, 1), bias=bias),
            nn.BatchNorm2d(outc), nn.SiLU())
        self.p_conv = nn.Conv2d(inc, 2 * num_param, kernel_size=3, padding=1, stride=stride)
        nn.init.constant_(self.p_conv.weight, 0)
        base = round(math.sqrt(num_param))
        rows = [i // base for i in range(num_param)]
        cols = [i % base for i in range(num_param)]
        self.register_buffer("p_n", torch.tensor(rows + cols, dtype=torch.int64).view(1, 2 * num_param, 1, 1))

    def _sampling_grid(self, offset):
        """p = p_0 + p_n + offset with p_0 rebuilt on the host every call, like the reference."""
        n, h, w = self.num_param, offset.size(2), offset.size(3)
        ii, jj = torch.meshgrid(torch.arange(0, h * self.stride, self.stride),
                                torch.arange(0, w * self.stride, self.stride), indexing="ij")
        p0 = torch.cat([ii.reshape(1, 1, h, w).repeat(1, n, 1, 1), jj.reshape(1, 1, h, w).repeat(1, n, 1, 1)], 1)
        p0 = p0.to(device=offset.device, dtype=offset.dtype)
        return p0 + self.p_n + offset

    @staticmethod
    def _take(x, rows, cols):
        """x (B,C,H,W), rows/cols (B,h,w,N) int64 -> (B,C,h,w,N) through a channel-expanded flat index + gather."""
        b, c, hx, wx = x.shape
        _, h, w, n = rows.shape
        flat = (rows * wx + cols).unsqueeze(1).expand(-1, c, -1, -1, -1).contiguous().view(b, c, -1)
        return x.contiguous().view(b, c, hx * wx).gather(dim=-1, index=flat).view(b, c, h, w, n)

    def forward(self, x):
        n = self.num_param
        hx, wx = x.size(2), x.size(3)
        offset = self.p_conv(x)
        p = self._sampling_grid(offset).contiguous().permute(0, 2, 3, 1)        # (B,h,w,2N)
        fl = p.detach().floor()
        r0 = torch.clamp(fl[..., :n], 0, hx - 1).long()
        k0 = torch.clamp(fl[..., n:], 0, wx - 1).long()
        r1 = torch.clamp(fl[..., :n] + 1, 0, hx - 1).long()
        k1 = torch.clamp(fl[..., n:] + 1, 0, wx - 1).long()
        pr = torch.clamp(p[..., :n], 0, hx - 1)
        pk = torch.clamp(p[..., n:], 0, wx - 1)
        g_lt = (1 + (r0.type_as(p) - pr)) * (1 + (k0.type_as(p) - pk))
        g_rb = (1 - (r1.type_as(p) - pr)) * (1 - (k1.type_as(p) - pk))
        g_lb = (1 + (r0.type_as(p) - pr)) * (1 - (k1.type_as(p) - pk))
        g_rt = (1 - (r1.type_as(p) - pr)) * (1 + (k0.type_as(p) - pk))
        samp = g_lt.unsqueeze(1) * self._take(x, r0, k0) + g_rb.unsqueeze(1) * self._take(x, r1, k1) \
            + g_lb.unsqueeze(1) * self._take(x, r0, k1) + g_rt.unsqueeze(1) * self._take(x, r1, k0)
        b, c, h, w, _ = samp.shape
        stacked = samp.permute(0, 1, 2, 4, 3).reshape(b, c, h * n, w)           # 'b c h w n -> b c (h n) w'
        return self.conv(stacked)
