"""torch.profiler kernel summary of one DEAL-YOLO-LD training step (config 4 harness): where the step time goes.
    python benchmarks/profile_train.py [--batch 64]
"""
import argparse
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import dealyolo  # noqa: E402
from experiment_yolo_b200 import dist as xdist  # noqa: E402
from experiment_yolo_b200.loss import DealYoloLoss, synthetic_uav_targets  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    model = dealyolo.DealYolo(nc=6)
    model.load_state_dict(dealyolo.seeded_state(model, 0))
    model = dealyolo.channels_last_(model.to(dev)).train()
    B = args.batch
    x = torch.rand((B, 3, 640, 640), device=dev).contiguous(memory_format=torch.channels_last)
    crit = DealYoloLoss(nc=6, strides=[float(v) for v in model.stride], max_boxes=16).to(dev)
    batch = synthetic_uav_targets(B, boxes_per_image=16, nc=6, seed=200, device=dev)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=0.01, momentum=0.937, nesterov=True)
    red = xdist.FlatGradAllReduce(model.parameters())

    def step():
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            opt.zero_grad(set_to_none=True)
            outs = model(x)
        loss, _ = crit(outs, batch)
        loss.backward()
        red()
        torch.nn.utils.clip_grad_norm_(params, max_norm=10.0)      # reference optimizer_step, engine/trainer.py:952
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
    tot = sum(r.device_time_total for r in rows)
    print(f"total device time {tot / 1e3:.1f} ms, batch {B}")
    for r in rows[:32]:
        print(f"{r.device_time_total / 1e3:9.2f} ms {100 * r.device_time_total / tot:5.1f}% x{r.count:4d}  {r.key[:110]}")


if __name__ == "__main__":
    main()
