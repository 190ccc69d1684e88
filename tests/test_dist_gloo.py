"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing: batch sharding with no collective for inference, one flat
gradient all-reduce(sum) for training.  The CPU processes use the eager port of LDConv from oracle/ --
the CUDA module has no CPU path -- so what is covered here is the host-side logic, not the kernels."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from experiment_yolo_b200 import dist as xdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_bounds_cover_the_batch_without_overlap():
    for gb in (1, 7, 8, 64, 65, 128):
        for world in (1, 2, 3, 4, 8):
            spans = [xdist.shard_bounds(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from experiment_yolo_b200 import dealyolo
    from oracle.ldconv_torch_port import LDConvTorchPort
    try:
        model = dealyolo.DealYolo(nc=6, ldconv_cls=LDConvTorchPort)
        model.load_state_dict(dealyolo.seeded_state(model, 0))
        g = torch.Generator().manual_seed(5)
        images = torch.rand(4, 3, 64, 64, generator=g)
        # ---- inference: shard, run, gather only for the check (the product path has no collective) ----
        model.eval()
        with torch.no_grad():
            y_local, _ = model(xdist.shard_batch(images, rank, world))
            y_full, _ = model(images)
        lo, hi = xdist.shard_bounds(4, rank, world)
        assert torch.allclose(y_local, y_full[lo:hi], atol=1e-5)
        # ---- training: per-rank BN statistics, one flat all-reduce(sum) ----
        model.train()
        for m in model.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.eval()            # freeze statistics so that the sum over shards equals the full-batch gradient
        targets = [torch.zeros(hi - lo, 70, s, s) for s in (16, 8, 4)]
        red = xdist.FlatGradAllReduce(model.parameters())
        outs = model(images[lo:hi])
        xdist.surrogate_detection_loss(outs, targets).backward()
        assert all(red.launched), "every slice's all-reduce is launched from the gradient hooks, during backward"
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(red.params, red.views)), "gradients are views of the flat buffer"
        red()
        flat = red.flat.clone()
        # a second step on the same reducer: zero() resets the buffer (no double counting), same result
        red.zero()
        outs = model(images[lo:hi])
        xdist.surrogate_detection_loss(outs, targets).backward()
        red()
        assert torch.allclose(red.flat, flat, rtol=1e-5, atol=1e-7)
        if rank == 0:
            model.zero_grad()
            outs = model(images)
            xdist.surrogate_detection_loss(outs, [torch.zeros(4, 70, s, s) for s in (16, 8, 4)]).backward()
            ref = torch.cat([p.grad.reshape(-1) for p in red.params])
            rel = float((flat - ref).norm() / ref.norm())
            q.put(("ok", rel, red.numel))
        dist.barrier()
    except Exception as e:  # pragma: no cover
        if rank == 0:
            q.put(("err", repr(e), 0))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharding_and_flat_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    status, rel, numel = q.get(timeout=280)
    for p in procs:
        p.join(timeout=60)
    assert status == "ok", rel
    assert numel == 918304 - 16          # every trainable parameter (the frozen DFL conv is excluded) in ONE buffer
    assert rel <= 1e-4
    assert all(p.exitcode == 0 for p in procs)
