"""Detection loss on the GPU: the fused task-aligned assigner (csrc/ldconv_tal.cu through ldconv_tal_metric / ldconv_tal_assign)
against the plain-PyTorch statement of the same arithmetic (experiment_yolo_b200/loss.py, pinned against the reference by
tests/test_loss_cpu.py), and the whole criterion on the GPU against the reference-minted fixtures."""
import glob
import os

import numpy as np
import pytest
import torch

from experiment_yolo_b200.loss import DealYoloLoss, TaskAlignedAssigner, make_anchors, synthetic_uav_targets
from oracle.gen_loss_golden import make_feats

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "loss_*.npz")))


def _assigner_inputs(b, img, seed, boxes_per_image, dev):
    crit = DealYoloLoss(nc=6)
    batch = synthetic_uav_targets(b, boxes_per_image=boxes_per_image, seed=seed)
    if boxes_per_image >= 8:      # a few large, overlapping boxes so that anchors are claimed by several gts
        batch["bboxes"][::5, 2:] = 0.4
    feats = make_feats(b, img, seed + 1)
    anc, st = make_anchors([tuple(f.shape[2:]) for f in feats], crit.strides)
    x = torch.cat([f.reshape(b, crit.no, -1) for f in feats], 2)
    dist, scores = x.split((64, 6), 1)
    dist = dist.permute(0, 2, 1).reshape(b, -1, 4, 16).softmax(3).matmul(torch.arange(16.0))
    boxes = torch.cat((anc - dist[..., :2], anc + dist[..., 2:]), -1) * st
    tg = crit.preprocess(batch, b, (img, img), "cpu")
    labels, gtb = tg.split((1, 4), 2)
    mask = (gtb.sum(2, keepdim=True) > 0).float()
    return [t.to(dev) for t in (scores.permute(0, 2, 1).sigmoid().contiguous(), boxes, anc * st, labels, gtb, mask)]


@pytest.mark.parametrize("b,img,n", [(3, 256, 16), (2, 128, 5), (4, 640, 16)])
def test_fused_assigner_matches_torch_statement(b, img, n):
    dev = torch.device("cuda", 0)
    args = _assigner_inputs(b, img, 11 * b + n, n, dev)
    ref = TaskAlignedAssigner(10, 6, fused=False)(*args)
    got = TaskAlignedAssigner(10, 6, fused=True)(*args)
    assert bool(ref[3].any())
    assert torch.equal(ref[3], got[3])                                    # foreground mask
    assert torch.equal(ref[4][ref[3]], got[4][ref[3]])                    # assigned gt of every positive
    assert torch.equal(ref[0][ref[3]], got[0][ref[3]]) and torch.equal(ref[1][ref[3]], got[1][ref[3]])
    np.testing.assert_allclose(got[2].cpu().numpy(), ref[2].cpu().numpy(), rtol=2e-4, atol=1e-6)   # atanf / powf vs ATen


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[5:-4])
def test_loss_on_gpu_matches_reference_fixture(path):
    dev = torch.device("cuda", 0)
    z = np.load(path)
    img, calls, ci = int(z["img"]), int(z["calls"]), int(z["case_index"])
    batch = {k: torch.from_numpy(z[k]).to(dev) for k in ("batch_idx", "cls", "bboxes")}
    crit = DealYoloLoss(nc=6).to(dev)
    for c in range(calls):
        b = z[f"grad{c}_0"].shape[0]
        feats = [f.to(dev).requires_grad_(True) for f in make_feats(b, img, 2000 + 10 * ci + c)]
        total, items = crit(feats, batch)
        total.backward()
        assert abs(float(total.detach()) - float(z[f"total{c}"])) <= 2e-4 * abs(float(z[f"total{c}"]))
        np.testing.assert_allclose(items.cpu().numpy(), z[f"items{c}"], rtol=2e-4, atol=1e-5)
        assert abs(float(crit.wiou_loss.iou_mean) - float(z[f"iou_mean{c}"])) <= 1e-5
        for i, f in enumerate(feats):
            ref = z[f"grad{c}_{i}"]
            got = f.grad.cpu().numpy() if f.grad is not None else np.zeros_like(ref)
            rel = float(np.linalg.norm((got - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-30))
            assert rel <= 1e-3, (c, i, rel)
