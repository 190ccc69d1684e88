"""Per-kernel device time of the fused DEAL-YOLO-LD inference step as it really runs: eager launches back to back on one stream
(warm L2, programmatic dependent launch), measured by torch.profiler (CUPTI) -- the complement of the serialised, cold-cache ncu
launch list of benchmarks/profile_step.py.  Prints per-kernel totals over `--steps` steps and, with --per-launch, every launch of
one step in order.
    python benchmarks/profile_step_insitu.py [--batch 64] [--steps 5] [--per-launch]
"""
import argparse
import collections
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import dealyolo, engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--per-launch", action="store_true")
    ap.add_argument("--l0-variant", type=int, default=0, help="first-layer kernel: 0 = tensor cores, 1 = CUDA-core rows kernel")
    args = ap.parse_args()
    from experiment_yolo_b200 import _lib as _l
    _l.load().ldconv_debug_l0_variant(args.l0_variant)
    dev = torch.device("cuda", 0)
    model = dealyolo.DealYolo(nc=6)
    model.load_state_dict(dealyolo.seeded_state(model, 0))
    model = dealyolo.channels_last_(model.to(dev).bfloat16().eval())
    run = engine.FusedDealYolo(model)
    xs = [torch.rand(args.batch, 3, 640, 640, device=dev).bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(2)]
    with torch.inference_mode():
        for i in range(4):
            run(xs[i & 1])
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(args.steps):
                run(xs[i & 1])
            torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time > 0]
    ev.sort(key=lambda e: e.time_range.start)
    short = lambda n: re.sub(r"[<(].*", "", n.replace("void ", "").replace("ldc::", ""))[:44]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in ev:
        a = agg[short(e.name)]
        a[0] += 1
        a[1] += e.device_time
    tot = sum(v[1] for v in agg.values())
    span = (ev[-1].time_range.end - ev[0].time_range.start) / args.steps
    print(f"{len(ev) // args.steps} launches per step, sum of kernel times {tot / args.steps:.1f} us per step, wall span {span:.1f} us per step")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"{n:46s} x{c // args.steps:3d} {t / args.steps:9.1f} us {100 * t / tot:5.1f}%")
    if args.per_launch:
        per = len(ev) // args.steps
        for e in ev[-per:]:
            print(f"  {e.device_time:8.1f} us  {short(e.name)}")


if __name__ == "__main__":
    main()
