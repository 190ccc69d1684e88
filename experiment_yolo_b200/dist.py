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
    """ONE flat fp32 gradient buffer for the whole model, reduced with all-reduce(sum).

    The parameters' `.grad` tensors ARE views of the flat buffer, so there is no pack / unpack copy per parameter: autograd
    accumulates in place, `zero()` is one memset and the optimizer reads the reduced values where they lie.  The buffer is
    cut into `buckets` contiguous slices in reverse parameter order (backward produces the last layers' gradients first); a
    post-accumulate hook counts the gradients of a slice and launches its all-reduce asynchronously the moment the last one
    has landed -- i.e. right after that slice's last weight-gradient kernel, overlapped with the rest of backward.  `__call__()`
    after `backward()` launches whatever has not been launched (parameters that received no gradient) and waits.
    Per-GPU BatchNorm statistics and plain sum like the reference's DDP step (engine/trainer.py:695,803-804)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], buckets: int = 2):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.views, ofs = [], 0
        for p in self.params:
            n = p.numel()
            chunk = self.flat[ofs:ofs + n]
            # same strides as the parameter (autograd's gradient layout contract): channels_last weights get a channels_last view
            # of their slice, so AccumulateGrad adds in place without a layout-converting copy
            dense_cl = p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last)
            self.views.append(chunk.as_strided(p.size(), p.stride()) if dense_cl else chunk.view_as(p))
            ofs += n
        # bucket k covers parameters [cut[k], cut[k+1]); bucket len-1 (the last layers) is complete first during backward
        buckets = max(1, min(int(buckets), len(self.params) or 1))
        target, self.cut, acc = self.numel / buckets, [0], 0
        for i, p in enumerate(self.params):
            acc += p.numel()
            if acc >= target * len(self.cut) and len(self.cut) < buckets and i + 1 < len(self.params):
                self.cut.append(i + 1)
        self.cut.append(len(self.params))
        self.bucket_of = {}
        for k in range(len(self.cut) - 1):
            for i in range(self.cut[k], self.cut[k + 1]):
                self.bucket_of[id(self.params[i])] = (k, i)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.works, self.launched, self.pending = [], [], []
        self.zero()

    # ---- per step -----------------------------------------------------------------------------------------------------------
    def zero(self):
        """replaces optimizer.zero_grad(): one memset, gradients stay views of the flat buffer"""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v
        self.pending = [self.cut[k + 1] - self.cut[k] for k in range(len(self.cut) - 1)]
        self.launched = [False] * (len(self.cut) - 1)
        self.works = []

    def _on_grad(self, p: torch.nn.Parameter):
        k, i = self.bucket_of[id(p)]
        v = self.views[i]
        if p.grad is not None and p.grad.data_ptr() != v.data_ptr():      # someone reset .grad (zero_grad(set_to_none=True)): adopt
            v.copy_(p.grad)
            p.grad = v
        self.pending[k] -= 1
        if self.pending[k] == 0:
            self._launch(k)

    def _slice(self, k: int) -> torch.Tensor:
        lo = sum(p.numel() for p in self.params[: self.cut[k]])
        hi = lo + sum(p.numel() for p in self.params[self.cut[k]: self.cut[k + 1]])
        return self.flat[lo:hi]

    def _launch(self, k: int):
        if self.launched[k]:
            return
        self.launched[k] = True
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.works.append(dist.all_reduce(self._slice(k), op=dist.ReduceOp.SUM, async_op=True))

    def __call__(self):
        """after backward(): reduce what is still local, wait for every slice; `.grad` then holds the summed gradients"""
        for p, v in zip(self.params, self.views):      # parameters autograd never touched keep their zero view
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
        for k in range(len(self.cut) - 1):
            self._launch(k)
        for w in self.works:
            w.wait()
        self.works = []
        return None

    def unpack(self):      # kept for callers of the first version: gradients already are views of the buffer
        return None


def surrogate_detection_loss(outs: List[torch.Tensor], targets: List[torch.Tensor]) -> torch.Tensor:
    """Stand-in for the reference's v8DetectionLoss in the synthetic training-step harness: a dense regression of the raw
    head maps (train-mode Detect output, nn/modules/head.py:47-48) onto synthetic targets, summed over the local batch
    like the reference loss (`loss.sum() * batch_size`, utils/loss.py:361).  The reference criterion itself (TAL + Wise-IoU +
    NWD + DFL) is experiment_yolo_b200/loss.py; this stand-in stays for A/B runs and the gloo test."""
    total = outs[0].new_zeros((), dtype=torch.float32)
    for o, t in zip(outs, targets):
        total = total + (o.float() - t.float()).square().mean(dim=(1, 2, 3)).sum()
    return total


def train_step(model: torch.nn.Module, images: torch.Tensor, targets: List[torch.Tensor],
               optimizer: torch.optim.Optimizer, reducer: FlatGradAllReduce) -> torch.Tensor:
    """One data-parallel step on this rank's shard: forward (train-mode BN, per-GPU statistics), backward through the
    CUDA LDConv kernels, ONE flat gradient all-reduce, optimizer step."""
    reducer.zero()
    outs = model(images)
    loss = surrogate_detection_loss(outs, targets)
    loss.backward()
    reducer()
    optimizer.step()
    return loss.detach()
