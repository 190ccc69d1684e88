"""Where the detection loss (experiment_yolo_b200/loss.py) spends its device time at the config-4 shape: batch 128, 640x640,
16 boxes per image.  Prints forward+backward ms and the top CUDA kernels by total time (torch.profiler).
    python benchmarks/profile_loss.py [--batch 128]
"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200.loss import DealYoloLoss, synthetic_uav_targets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--img", type=int, default=640)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    B = args.batch
    g = torch.Generator(device=dev).manual_seed(0)
    feats = [torch.randn((B, 70, args.img // s, args.img // s), device=dev, generator=g).bfloat16().requires_grad_(True) for s in (4, 8, 16)]
    crit = DealYoloLoss(nc=6, max_boxes=16).to(dev)
    batch = synthetic_uav_targets(B, 16, 6, seed=1, device=dev)

    def step():
        for f in feats:
            f.grad = None
        loss, _ = crit(feats, batch)
        loss.backward()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("loss fwd+bwd ms:", e0.elapsed_time(e1) / 5, "peak MB:", torch.cuda.max_memory_allocated() / 1e6)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))


if __name__ == "__main__":
    main()
