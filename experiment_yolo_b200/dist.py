"""Multi-GPU plumbing of the hot path: one process per GPU, `torch.distributed` for rendezvous.

* Inference is embarrassingly parallel over the batch (every LDConv sample depends on its own image only): the batch is
  sharded across ranks, the 3.7 MB of weights are replicated, and there is NO data-path collective.
* Training is data parallel with per-GPU BatchNorm statistics, like the reference's plain DDP without SyncBN
  (/root/reference/ultralytics/engine/trainer.py:695): all gradients (918,304 fp32 = 3.67 MB, LDConv share 80,798) are
  packed into ONE flat buffer and reduced with a single all-reduce(sum) per step -- latency-bound at this size, so one
  call instead of DDP's bucket hooks.  The reference's loss scaling (loss * world_size under DDP averaging,
  trainer.py:803-804) nets out to a plain SUM of per-rank gradients, which is what is done here.
The code is backend-agnostic (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of a global batch; the first `global_batch % world` ranks get one extra."""
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


class FlatGradAllReduce:
    """Packs the gradients of `params` into one flat fp32 buffer, all-reduces it once (sum) and scatters it back."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)

    def __call__(self, async_op: bool = False):
        ofs = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[ofs:ofs + n].zero_()
            else:
                self.flat[ofs:ofs + n].copy_(p.grad.reshape(-1))
            ofs += n
        work = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        if async_op and work is not None:
            return work
        self.unpack()
        return None

    def unpack(self):
        ofs = 0
        for p in self.params:
            n = p.numel()
            g = self.flat[ofs:ofs + n].view_as(p)
            if p.grad is None:
                p.grad = g.to(p.dtype).clone()
            else:
                p.grad.copy_(g)
            ofs += n


def surrogate_detection_loss(outs: List[torch.Tensor], targets: List[torch.Tensor]) -> torch.Tensor:
    """Stand-in for the reference's v8DetectionLoss in the synthetic training-step harness: a dense regression of the raw
    head maps (train-mode Detect output, nn/modules/head.py:47-48) onto synthetic targets, summed over the local batch
    like the reference loss (`loss.sum() * batch_size`, utils/loss.py:361).  The TAL / WIoU / NWD loss itself is a
    "next" row and is not re-implemented."""
    total = outs[0].new_zeros((), dtype=torch.float32)
    for o, t in zip(outs, targets):
        total = total + (o.float() - t.float()).square().mean(dim=(1, 2, 3)).sum()
    return total


def train_step(model: torch.nn.Module, images: torch.Tensor, targets: List[torch.Tensor],
               optimizer: torch.optim.Optimizer, reducer: FlatGradAllReduce) -> torch.Tensor:
    """One data-parallel step on this rank's shard: forward (train-mode BN, per-GPU statistics), backward through the
    CUDA LDConv kernels, ONE flat gradient all-reduce, optimizer step."""
    optimizer.zero_grad(set_to_none=True)
    outs = model(images)
    loss = surrogate_detection_loss(outs, targets)
    loss.backward()
    reducer()
    optimizer.step()
    return loss.detach()
