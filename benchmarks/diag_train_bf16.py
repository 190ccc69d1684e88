"""Diagnostic: head maps of one bf16-autocast TRAINING forward against the reference fixture (tests/golden/train_step.npz), with the
library's training paths switched on / off one by one."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import dealyolo  # noqa: E402

z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "train_step.npz"))
dev = "cuda:0"


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))


def run(train, autocast, tag):
    m = dealyolo.DealYolo(nc=6)
    m.load_state_dict(dealyolo.seeded_state(m, seed=0), strict=True)
    m = dealyolo.channels_last_(m.to(dev))
    m.train() if train else m.eval()
    x = torch.from_numpy(z["x"]).to(dev).contiguous(memory_format=torch.channels_last)
    feats_by_layer = {}
    hooks = [mod.register_forward_hook(lambda mod, i, o, k=k: feats_by_layer.__setitem__(k, o.detach().float().cpu().numpy()) if torch.is_tensor(o) else None)
             for k, mod in enumerate(m.model)]
    with torch.no_grad() if not train else torch.enable_grad():
        if autocast:
            with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
                out = m(x)
        else:
            out = m(x)
    for h in hooks:
        h.remove()
    feats = out if train else out[1]
    print(tag, [round(rel(f.detach().float().cpu().numpy(), z[f"feat{i}"]), 4) for i, f in enumerate(feats)], flush=True)
    return feats_by_layer


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ref32 = run(True, False, "fp32 train            ")
a = run(True, True, "bf16 train, all library")
print("  per-layer rel-L2 vs the fp32 GPU run:", {k: round(rel(a[k], ref32[k]), 3) for k in sorted(a) if k in ref32})
dealyolo.Conv.fused_bn_silu_train = False
b = run(True, True, "bf16 train, torch BN   ")
dealyolo.ScalSeq.fused_train_tail = False
c = run(True, True, "bf16 train, torch SSFF ")
print("  per-layer rel-L2 vs the fp32 GPU run:", {k: round(rel(c[k], ref32[k]), 3) for k in sorted(c) if k in ref32})
