"""Mint golden vectors for the TRAINING-mode SSFF block FROM THE REFERENCE's own classes (authoring container only).

Imports the unmodified reference (`/root/reference/ultralytics`, stub importer as in gen_model_golden.py), builds
`ScalSeq([32, 64, 128], 32)` and `Add()` (nn/extra_modules/block.py:3414-3443, 3479-3484), loads seeded parameters, runs one
TRAINING forward (batch statistics, running-statistics update) + backward in fp32 on the CPU and stores inputs, parameters (the
state_dict, whose keys the benchmark graph's ScalSeq must share), output, every gradient and the updated running statistics under
tests/golden/ssff_train.npz.  tests/test_ssff_golden.py holds experiment_yolo_b200.dealyolo.ScalSeq / Add to it on the CPU, and
(on a B200) the library's training tail (train_ops.scalseq_tail) on bf16-rounded tensors.

    python oracle/gen_ssff_golden.py

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle.gen_model_golden import load_reference_tasks  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "ssff_train.npz")


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(4)
    load_reference_tasks()
    from ultralytics.nn.extra_modules.block import Add, ScalSeq
    torch.manual_seed(7)
    ref = ScalSeq([32, 64, 128], 32).train()
    with torch.no_grad():
        for m in ref.modules():
            if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.2)
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
        ref.conv3d.bias.normal_(0, 0.5)
    sd0 = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    xs = [torch.randn(2, c, h, w) for c, h, w in ((32, 16, 24), (64, 8, 12), (128, 4, 6))]
    extra = torch.randn(2, 32, 16, 24)
    ins = [t.clone().requires_grad_(True) for t in xs]
    ex = extra.clone().requires_grad_(True)
    y = Add()([ex, ref(ins)])
    gout = torch.randn(y.shape)
    y.backward(gout)
    out = {"out": y.detach().numpy(), "grad_out": gout.numpy(), "extra": extra.numpy(), "grad_extra": ex.grad.numpy()}
    for i, (t, g) in enumerate(zip(xs, ins)):
        out[f"x{i}"] = t.numpy()
        out[f"grad_x{i}"] = g.grad.numpy()
    for k, v in sd0.items():
        out["param." + k] = v.numpy()
    for k, p in ref.named_parameters():
        out["grad." + k] = p.grad.numpy()
    for k, v in ref.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["after." + k] = v.detach().numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(out), "arrays; keys e.g.", list(sd0.keys())[:4])


if __name__ == "__main__":
    main()
