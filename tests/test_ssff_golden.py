"""SSFF (`ScalSeq` + `Add`, reference nn/extra_modules/block.py:3414-3443, 3479-3484) in TRAINING mode against a fixture minted from the
reference's own classes (oracle/gen_ssff_golden.py -> tests/golden/ssff_train.npz): forward with batch statistics, every gradient,
the running-statistics update.  CPU: the benchmark graph's modules in fp32.  GPU: the same modules under bf16 autocast, whose tail
runs through the library (train_ops.scalseq_tail, ldconv_add_nhwc, the BatchNorm / SiLU passes of the 1x1 Conv blocks)."""
import os

import numpy as np
import pytest
import torch

from experiment_yolo_b200 import dealyolo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ssff_train.npz")


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def _build(z, device):
    seq = dealyolo.ScalSeq([32, 64, 128], 32)
    sd = {k[len("param."):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param.")}
    assert list(seq.state_dict().keys()) == list(sd.keys())                 # the reference's parameter layout, key for key
    seq.load_state_dict(sd, strict=True)
    return seq.to(device).train()


def _run(z, device, autocast):
    seq = _build(z, device)
    ins = [torch.from_numpy(z[f"x{i}"]).to(device).contiguous(memory_format=torch.channels_last).requires_grad_(True) for i in range(3)]
    ex = torch.from_numpy(z["extra"]).to(device).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    if autocast:
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            y = dealyolo.Add()([ex, seq(ins)])
    else:
        y = dealyolo.Add()([ex, seq(ins)])
    y.float().backward(torch.from_numpy(z["grad_out"]).to(device))
    return seq, ins, ex, y


def test_scalseq_add_training_fp32_cpu_matches_reference_fixture():
    z = np.load(GOLD)
    seq, ins, ex, y = _run(z, "cpu", False)
    assert np.abs(y.detach().numpy() - z["out"]).max() <= 1e-5
    for i in range(3):
        assert _rel(ins[i].grad.numpy(), z[f"grad_x{i}"]) <= 1e-4
    assert _rel(ex.grad.numpy(), z["grad_extra"]) <= 1e-6
    for k, p in seq.named_parameters():
        ref = z["grad." + k]
        if k == "conv3d.bias":          # BatchNorm makes it exactly zero in theory (fp32 cancellation noise ~1e-4 in both): absolute
            assert np.abs(p.grad.numpy() - ref).max() <= 1e-3 and np.abs(ref).max() <= 1e-3
        else:
            assert _rel(p.grad.numpy(), ref) <= 2e-4, k
    for k, v in seq.state_dict().items():
        if "running" in k:
            assert np.abs(v.numpy() - z["after." + k]).max() <= 1e-5, k
        if "num_batches" in k:
            assert int(v) == int(z["after." + k])


@pytest.mark.gpu
def test_scalseq_add_training_bf16_library_tail_matches_reference_fixture():
    """bf16 autocast on the GPU: the tail goes through the library (checked by the call counters); tolerance = bf16 storage of
    every intermediate against the reference's fp32 run (outputs rel-L2 <= 2e-2, gradients <= 6e-2, running statistics <= 1e-2)."""
    from experiment_yolo_b200 import _lib
    z = np.load(GOLD)
    _lib.call_counts.clear()
    seq, ins, ex, y = _run(z, "cuda:0", True)
    for name in ("ldconv_ssff_max_fwd", "ldconv_ssff_max_bwd", "ldconv_add_nhwc", "ldconv_upsample_nearest", "ldconv_upsample_nearest_bwd",
                 "ldconv_col_stats"):
        assert name in _lib.call_counts, name
    assert _rel(y.float().detach().cpu().numpy(), z["out"]) <= 2e-2
    for i in range(3):
        assert _rel(ins[i].grad.float().cpu().numpy(), z[f"grad_x{i}"]) <= 6e-2, i
    assert _rel(ex.grad.float().cpu().numpy(), z["grad_extra"]) <= 1e-2
    for k, p in seq.named_parameters():
        if k == "conv3d.bias":
            continue
        assert _rel(p.grad.float().cpu().numpy(), z["grad." + k]) <= 6e-2, k
    for k, v in seq.state_dict().items():
        if "running" in k:
            assert _rel(v.float().cpu().numpy(), z["after." + k]) <= 1e-2, k
