"""Host-link roofline of the end-to-end path: every rank copies the benchmark's input batch (64 x 3 x 640 x 640 uint8 = 78.6 MB) from
pinned host memory to its GPU in a loop, all ranks at once (one cudaMemcpyAsync per copy, two alternating host buffers); rank 0
prints the aggregate GB/s and the images/s that bandwidth would carry.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 benchmarks/h2d_ceiling.py"""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    host = [torch.randint(0, 256, (64, 3, 640, 640), dtype=torch.uint8).pin_memory() for _ in range(2)]
    dst = [torch.empty((64, 3, 640, 640), device=dev, dtype=torch.uint8) for _ in range(2)]
    for i in range(5):
        dst[i & 1].copy_(host[i & 1], non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        dst[i & 1].copy_(host[i & 1], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(t.item())
        nbytes = host[0].numel()
        gbs = world * n * nbytes / ms / 1e6
        print(json.dumps({"n_gpus": world, "h2d_aggregate_GBps": round(gbs, 1), "per_gpu_GBps": round(gbs / world, 1),
                          "images_per_s_ceiling": round(world * n * 64 / ms * 1e3, 0), "copy_MB": round(nbytes / 1e6, 1), "copies_per_rank": n}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
