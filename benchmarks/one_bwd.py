"""One LDConv training forward + backward through the module at a yolov8-LD-P2 layer shape, between cudaProfilerStart/Stop
(for `ncu --profile-from-start off` launch lists of the backward path).
    python benchmarks/one_bwd.py --layer 1 [--batch 64] [--fp32]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import experiment_yolo_b200 as E  # noqa: E402
from benchmarks.ldconv_layers import LAYERS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layer", type=int, default=1)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--fp32", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    li, C, O, N, s, H = [l for l in LAYERS if l[0] == args.layer][0]
    dtype = torch.float32 if args.fp32 else torch.bfloat16
    torch.manual_seed(0)
    mod = E.LDConv(C, O, N, s).to(dev)
    with torch.no_grad():
        mod.p_conv.weight.normal_(0, 0.05)
    mod = mod.to(dtype).train()
    x = torch.randn(args.batch, C, H, H, device=dev).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    h = (H - 1) // s + 1
    gout = torch.randn(args.batch, O, h, h, device=dev).to(dtype).contiguous(memory_format=torch.channels_last)
    for _ in range(2):
        mod.zero_grad(set_to_none=True)
        x.grad = None
        mod(x).backward(gout)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    mod.zero_grad(set_to_none=True)
    x.grad = None
    mod(x).backward(gout)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled one fwd+bwd, layer", li)


if __name__ == "__main__":
    main()
