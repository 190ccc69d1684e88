"""BASELINE.json config 4: DEAL-YOLO-LD training step on synthetic 640x640 batches, random-init weights, bf16 autocast,
per-GPU BatchNorm statistics, ONE flat gradient all-reduce over NCCL per step (experiment_yolo_b200/dist.py), SGD with
nesterov momentum (lr 0.01, momentum 0.937: reference engine/trainer.py:1164, cfg/default.yaml:90-92).
The loss is the reference's criterion for this config -- task-aligned assignment + BCE + Wise-IoU v3 + NWD + DFL
(experiment_yolo_b200/loss.py, pinned against the reference's v8DetectionLoss by tests/test_loss_cpu.py) -- on synthetic UAV
targets (16 boxes of 6-32 px per image); --loss surrogate keeps the earlier dense stand-in for A/B.
    python benchmarks/train_step.py [--batch 128] [--steps 10]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/train_step.py --batch 128
(--batch is the GLOBAL batch, split evenly over ranks.)
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from experiment_yolo_b200 import _lib, dealyolo  # noqa: E402
from experiment_yolo_b200 import dist as xdist  # noqa: E402
from experiment_yolo_b200.loss import DealYoloLoss, synthetic_uav_targets  # noqa: E402


def run_train(global_batch: int, steps: int, warmup: int, rank: int, world: int, dev, img: int = 640, loss_kind: str = "wiou_nwd"):
    """BASELINE config 4 on an initialised process group (or a single rank); returns the result dict on every rank."""
    torch.backends.cudnn.benchmark = True
    model = dealyolo.DealYolo(nc=6)
    model.load_state_dict(dealyolo.seeded_state(model, 0))
    model = dealyolo.channels_last_(model.to(dev)).train()
    lo, hi = xdist.shard_bounds(global_batch, rank, world)
    B = hi - lo
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.rand((B, 3, img, img), device=dev, generator=g).contiguous(memory_format=torch.channels_last)
    targets = [torch.zeros((B, 70, img // s, img // s), device=dev) for s in (4, 8, 16)]
    crit = DealYoloLoss(nc=6, strides=[float(v) for v in model.stride], max_boxes=16).to(dev)
    batch = synthetic_uav_targets(B, boxes_per_image=16, nc=6, seed=200 + rank, device=dev)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=0.01, momentum=0.937, nesterov=True)
    red = xdist.FlatGradAllReduce(model.parameters())

    def step():
        with torch.autocast(device_type="cuda", dtype=torch.bfloat16):
            red.zero()      # gradients are views of the flat all-reduce buffer: one memset instead of ~200 tensors
            outs = model(x)
        if loss_kind == "surrogate":
            loss = xdist.surrogate_detection_loss(outs, targets)
        else:       # loss.sum() * local batch like the reference (utils/loss.py:361); the all-reduce sums over ranks
            loss, _ = crit(outs, batch)
        loss.backward()      # the slices of the flat buffer are all-reduced from the gradient hooks while backward still runs
        red()
        torch.nn.utils.clip_grad_norm_(params, max_norm=10.0)      # reference optimizer_step, engine/trainer.py:952
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    _lib.call_counts.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    return {"metric": "train_images_per_sec", "value": round(global_batch / ms * 1e3, 1), "unit": "images/s", "ms_per_step": round(ms, 3),
            "n_gpus": world, "global_batch": global_batch, "per_gpu_batch": B, "steps": steps, "scaling": "strong",
            "dtype": "bf16 autocast (fp32 master weights)",
            "loss": ("dense surrogate (not the reference WIoU+NWD loss)" if loss_kind == "surrogate" else
                     "TAL(topk 10) + BCE + Wise-IoU v3 + NWD (ratio 0.5) + DFL, box 7.5 / cls 0.5 / dfl 1.5, 16 synthetic UAV boxes per image"),
            "final_loss": float(loss.detach()),
            "grad_allreduce": f"one flat fp32 buffer of {red.numel} values in {len(red.cut) - 1} slices, each all-reduced (sum) from a "
                              "gradient hook during backward" + (" over NCCL" if world > 1 else " (single rank: no collective)"),
            "ldconv_calls_per_step": {k: v // steps for k, v in sorted(_lib.call_counts.items())}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--img", type=int, default=640)
    ap.add_argument("--loss", default="wiou_nwd", choices=["wiou_nwd", "surrogate"])
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    res = run_train(args.batch, args.steps, args.warmup, rank, world, dev, args.img, args.loss)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
