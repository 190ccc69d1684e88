"""One LDConv layer (C = O) at a sweep shape: which C-ABI entry points the inference forward calls and a torch.profiler table
of its kernels.    python scripts/one_cfg.py C N H s B"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import experiment_yolo_b200 as E  # noqa: E402
from experiment_yolo_b200 import _lib  # noqa: E402

C, N, H, s, B = [int(v) for v in sys.argv[1:6]]
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mod = E.LDConv(C, C, N, s).to(dev)
with torch.no_grad():
    mod.p_conv.weight.normal_(0, 0.05)
mod = mod.bfloat16().eval()
x = torch.randn(B, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
_lib.call_counts.clear()
with torch.no_grad():
    for _ in range(3):
        y = mod(x)
    torch.cuda.synchronize()
    print("ok", sorted(_lib.call_counts), float(y.float().abs().mean()))
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            mod(x)
        torch.cuda.synchronize()
for r in sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:8]:
    print(f"{r.device_time_total / r.count:9.1f} us x{r.count:3d}  {r.key[:120]}")
