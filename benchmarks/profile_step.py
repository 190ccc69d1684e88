"""One eager DEAL-YOLO-LD forward (batch 64, 640x640, bf16, channels_last) between cudaProfilerStart/Stop, for
`ncu --profile-from-start off` launch lists (every kernel of exactly one step, after cuDNN autotuning has settled).
    python benchmarks/profile_step.py [--batch 64] [--fp32]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import dealyolo, engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--img", type=int, default=640)
    ap.add_argument("--engine", default="fused", choices=["fused", "eager"])
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    model = dealyolo.DealYolo(nc=6)
    model.load_state_dict(dealyolo.seeded_state(model, 0))
    model = dealyolo.channels_last_(model.to(dev).bfloat16().eval())
    run = engine.FusedDealYolo(model) if args.engine == "fused" else model
    x = torch.rand(args.batch, 3, args.img, args.img, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
    with torch.inference_mode():
        for _ in range(3):
            run(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        run(x)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    print("profiled one step")


if __name__ == "__main__":
    main()
