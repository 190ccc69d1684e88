"""Mint golden vectors for the LDConv hot path FROM THE REFERENCE ITSELF.

Runs only in the authoring container (needs /root/reference, which is read-only and absent on the GPU box).  It
imports the reference's ultralytics/nn/modules/conv.py unmodified (the file needs only torch, numpy, einops;
SURVEY.md Appendix D), runs `LDConv` (conv.py:350-503) on seeded inputs in fp32 on the CPU and stores inputs,
parameters, intermediates, outputs and autograd gradients as small .npz fixtures under tests/golden/.

    python oracle/gen_golden.py            # rewrites tests/golden/ldconv_*.npz and tests/golden/MANIFEST.json

The reference has no tests or golden vectors of its own for this path; these fixtures are what
pins oracle/ldconv_oracle.c and, through it, the CUDA kernels.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import os
import sys
import warnings

import numpy as np
import torch

REF_CONV = "/root/reference/ultralytics/nn/modules/conv.py"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# name, inc, outc, num_param, stride, B, H, W, p_conv weight sigma, p_conv bias sigma (None = reference default init),
# bn eps, bn momentum
CASES = [
    # the YAML's own layer shapes, shrunk spatially (cfg/models/yolov8-LD-P2.yaml:15-44)
    ("l0_3to16_n3s2",    3, 16, 3, 2, 2, 16, 20, 0.05, None, 1e-3, 0.03),
    ("l1_16to32_n3s2",  16, 32, 3, 2, 1, 12, 10, 0.05, None, 1e-3, 0.03),
    ("l8_16to8_n1s1",   16,  8, 1, 1, 2,  7,  9, 0.05, None, 1e-3, 0.03),
    # odd sizes, every num_param family of _get_p_n (conv.py:413-432), both strides
    ("n2s1_odd",         4,  8, 2, 1, 2, 13, 17, 0.10, 2.0, 1e-5, 0.1),
    ("n4s2_odd",         4,  8, 4, 2, 1, 21, 33, 0.10, 2.0, 1e-5, 0.1),
    ("n5s1",             8,  8, 5, 1, 2,  9, 11, 0.30, 1.0, 1e-5, 0.1),
    ("n5s2_odd",         5,  6, 5, 2, 2, 13, 17, 0.10, 2.0, 1e-3, 0.03),
    ("n7s1",             4,  4, 7, 1, 1, 10,  8, 0.20, 2.0, 1e-5, 0.1),
    ("n9s2",             8, 16, 9, 2, 2, 16, 16, 0.05, 0.5, 1e-5, 0.1),
    ("n9s1_odd",         2,  4, 9, 1, 1, 11, 13, 0.30, 3.0, 1e-5, 0.1),
    # offsets far outside the image: exercises the independent clamp of corners and p (conv.py:379-393)
    ("n3s2_far",         4,  8, 3, 2, 2, 10, 14, 0.50, 8.0, 1e-5, 0.1),
    # as-shipped init: p_conv.weight == 0 (conv.py:357), offsets = bias only
    ("n3s2_shipped",     8,  8, 3, 2, 2, 12, 12, 0.00, None, 1e-3, 0.03),
    # exactly-zero offsets: last row / column doubling at stride 1, edge samples at stride 2 (SURVEY.md App. C 1,4)
    ("n1s1_zero",        1,  2, 1, 1, 1,  4,  4, 0.00, 0.0, 1e-5, 0.1),
    ("n3s2_zero",        1,  2, 3, 2, 1,  6,  6, 0.00, 0.0, 1e-5, 0.1),
]


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_conv", REF_CONV)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    return ref


def run_case(ref, name, inc, outc, N, s, B, H, W, w_sigma, b_sigma, eps, momentum, seed):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    m = ref.LDConv(inc, outc, N, s)
    bn = m.conv[1]
    bn.eps, bn.momentum = eps, momentum
    with torch.no_grad():
        if w_sigma > 0:
            m.p_conv.weight.copy_(torch.randn(m.p_conv.weight.shape, generator=g) * w_sigma)
        if b_sigma is not None:
            m.p_conv.bias.copy_(torch.randn(m.p_conv.bias.shape, generator=g) * b_sigma)
        m.conv[0].weight.copy_(torch.randn(m.conv[0].weight.shape, generator=g) * 0.3)
        bn.weight.copy_(torch.rand(outc, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(outc, generator=g) * 0.2)
        bn.running_mean.copy_(torch.randn(outc, generator=g) * 0.3)
        bn.running_var.copy_(torch.rand(outc, generator=g) + 0.5)
    if name.endswith("_zero"):
        x = torch.arange(1, B * inc * H * W + 1, dtype=torch.float32).view(B, inc, H, W)
    else:
        x = torch.randn(B, inc, H, W, generator=g)

    rec = {}
    params = {k: v.detach().clone().numpy() for k, v in m.state_dict().items()}

    # capture what the reference passes to its gathers and what it hands to the (N,1) conv
    qs = []
    orig_get_x_q = m._get_x_q
    m._get_x_q = lambda xx, q, n: (qs.append(q.detach().clone()), orig_get_x_q(xx, q, n))[1]
    captured = {}
    orig_reshape = ref.LDConv._reshape_x_offset
    m._reshape_x_offset = lambda xo, n: captured.setdefault("x_offset", orig_reshape(xo, n))

    # ---- eval forward -------------------------------------------------------------------------------------------
    m.eval()
    with torch.no_grad():
        out_eval = m(x)
        offset = m.p_conv(x)
        p = m._get_p(offset, offset.data.type()).contiguous().permute(0, 2, 3, 1)
        p = torch.cat([torch.clamp(p[..., :N], 0, H - 1), torch.clamp(p[..., N:], 0, W - 1)], dim=-1)
    q_lt, q_rb = qs[0], qs[1]                   # (B,h,w,2N) int64: rows then cols (conv.py:379-384)
    idx = torch.stack([q_lt[..., :N], q_rb[..., :N], q_lt[..., N:], q_rb[..., N:]], dim=-1).to(torch.int32)
    assert torch.equal(qs[2], torch.cat([q_lt[..., :N], q_rb[..., N:]], -1))   # q_lb = (lt row, rb col)
    assert torch.equal(qs[3], torch.cat([q_rb[..., :N], q_lt[..., N:]], -1))   # q_rt = (rb row, lt col)
    coord = torch.stack([p[..., :N], p[..., N:]], dim=-1)
    rec.update(x=x.numpy(), offset=offset.numpy(), idx=idx.numpy(), coord=coord.numpy(),
               x_offset=captured["x_offset"].detach().numpy(), out_eval=out_eval.numpy())

    # ---- eval backward (running statistics are constants) ---------------------------------------------------------
    qs.clear(); captured.clear()
    grad_out = torch.randn(out_eval.shape, generator=g)
    xe = x.clone().requires_grad_(True)
    m.zero_grad()
    m(xe).backward(grad_out)
    rec.update(grad_out=grad_out.numpy(), eval_grad_x=xe.grad.numpy(),
               eval_grad_conv0_weight=m.conv[0].weight.grad.numpy().copy(),
               eval_grad_p_conv_weight=m.p_conv.weight.grad.numpy().copy())

    # ---- train forward + backward -----------------------------------------------------------------------------------
    qs.clear(); captured.clear()
    m.train()
    m.zero_grad()
    xt = x.clone().requires_grad_(True)
    out_train = m(xt)
    out_train.backward(grad_out)
    rec.update(out_train=out_train.detach().numpy(), train_grad_x=xt.grad.numpy(),
               train_running_mean=bn.running_mean.numpy().copy(), train_running_var=bn.running_var.numpy().copy(),
               train_num_batches_tracked=bn.num_batches_tracked.numpy().copy())
    for k, prm in m.named_parameters():
        rec["train_grad_" + k.replace(".", "_")] = prm.grad.numpy().copy()
    for k, v in params.items():
        rec["param_" + k.replace(".", "_")] = v
    rec["meta"] = np.array([inc, outc, N, s, B, H, W], dtype=np.int64)
    rec["bn_cfg"] = np.array([eps, momentum], dtype=np.float64)
    return rec


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(1)
    torch.use_deterministic_algorithms(True)
    ref = load_reference()
    os.makedirs(OUT_DIR, exist_ok=True)
    manifest = {"generator": "oracle/gen_golden.py", "reference": REF_CONV + ":350-503",
                "torch": torch.__version__, "cases": {}}
    for ci, case in enumerate(CASES):
        rec = run_case(ref, *case, seed=1000 + ci)
        path = os.path.join(OUT_DIR, f"ldconv_{case[0]}.npz")
        np.savez_compressed(path, **rec)
        with open(path, "rb") as f:
            digest = hashlib.sha256(f.read()).hexdigest()[:16]
        manifest["cases"][case[0]] = {"args": list(case[1:]), "seed": 1000 + ci, "bytes": os.path.getsize(path),
                                      "sha256_16": digest}
        print(f"{case[0]:>18s}  {os.path.getsize(path) / 1024:7.1f} KiB")
    with open(os.path.join(OUT_DIR, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    sys.exit(main())
