"""Which aten::copy_ / aten::contiguous / aten::to calls the config-4 training step makes (shapes, counts, device time):
    python benchmarks/profile_train_copies.py [--batch 64]"""
import argparse
import collections
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
            red.zero()
            outs = model(x)
        loss, _ = crit(outs, batch)
        loss.backward()
        red()
        torch.nn.utils.clip_grad_norm_(params, max_norm=10.0)
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
        step()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0, ""])
    for e in prof.events():
        if e.name in ("aten::copy_",) and e.device_time_total > 0:
            # the chain of enclosing ops (backward copies have no Python stack: the autograd node's name is the attribution)
            chain, par = [], e.cpu_parent
            while par is not None and len(chain) < 6:
                chain.append(par.name.replace("autograd::engine::evaluate_function: ", "bwd:"))
                par = par.cpu_parent
            key = (str(e.input_shapes), " < ".join(chain)[-110:] if chain else "?")
            agg[key][0] += 1
            agg[key][1] += e.device_time_total
    rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
    tot = sum(v[1] for _, v in rows)
    print(f"aten::copy_ device time {tot / 1e3:.2f} ms in {sum(v[0] for _, v in rows)} calls")
    for (shapes, where), (n, t, _) in rows[:40]:
        print(f"{t / 1e3:7.3f} ms x{n:4d}  {shapes[:70]:70s}  {where}")


if __name__ == "__main__":
    main()
