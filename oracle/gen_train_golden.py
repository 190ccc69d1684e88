"""Mint a golden vector for ONE TRAINING STEP of BASELINE.json config 4 FROM THE REFERENCE (authoring container only).

Imports the unmodified reference (stub importer of oracle/gen_model_golden.py), builds `DetectionModel('yolov8-LD-P2.yaml', nc=6)` in
training mode with the seeded weights of experiment_yolo_b200.dealyolo.seeded_state, its `v8DetectionLoss` configured as
oracle/gen_loss_golden.py does (Wise-IoU v3 + NWD, ratio 0.5), runs forward -> criterion -> backward in fp32 on the CPU for a seeded
2 x 3 x 96 x 128 batch with seeded targets, and stores: the three head maps, the loss and its items, the gradient of EVERY parameter
(as its L2 norm, plus the full tensor for the LDConv rows and the first / last conv), and the BatchNorm running statistics after the
step (as norms).  tests/test_train_golden.py holds experiment_yolo_b200's DealYolo + DealYoloLoss to it.

    python oracle/gen_train_golden.py        # writes tests/golden/train_step.npz

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
import warnings
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle.gen_loss_golden import make_targets  # noqa: E402
from oracle.gen_model_golden import REF_YAML, load_reference_tasks  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "train_step.npz")
NC = 6
SPEC = [(9, 0.04, 0.3), (14, 0.03, 0.2)]
FULL = ("model.0.", "model.1.", "model.8.", "model.21.", "model.26.cv2.0.2.", "model.24.conv3d.")


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(8)
    T = load_reference_tasks()
    from ultralytics.utils.loss import v8DetectionLoss
    from ultralytics.utils.metrics import WiseIouLoss
    from experiment_yolo_b200 import dealyolo

    ref = T.DetectionModel(REF_YAML, ch=3, nc=NC, verbose=False)
    mine = dealyolo.DealYolo(nc=NC)
    ref.load_state_dict(dealyolo.seeded_state(mine, seed=0), strict=True)
    ref.args = SimpleNamespace(box=7.5, cls=0.5, dfl=1.5)
    ref.train()
    crit = v8DetectionLoss(ref)
    crit.bbox_loss.use_wiseiou = True
    crit.bbox_loss.wiou_loss = WiseIouLoss(ltype="WIoU", monotonous=False, inner_iou=False, focaler_iou=False)
    crit.bbox_loss.nwd_loss = True
    crit.bbox_loss.iou_ratio = 0.5
    x = torch.rand((2, 3, 96, 128), generator=torch.Generator().manual_seed(5))
    batch = make_targets(SPEC, 4242)
    feats = ref(x)
    total, items = crit(feats, batch)
    total.backward()
    rec = {"x": x.numpy(), "batch_idx": batch["batch_idx"].numpy(), "cls": batch["cls"].numpy(), "bboxes": batch["bboxes"].numpy(),
           "total": total.detach().numpy(), "items": items.numpy()}
    for i, f in enumerate(feats):
        rec[f"feat{i}"] = f.detach().numpy()
    names, norms = [], []
    for k, p in ref.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        names.append(k)
        norms.append(float(g.double().norm()))
        if k.startswith(FULL):
            rec["grad." + k] = g.numpy()
    rec["grad_names"] = np.array(names)
    rec["grad_norms"] = np.array(norms)
    bn_names, bn_norms = [], []
    for k, v in ref.state_dict().items():
        if "running_mean" in k or "running_var" in k:
            bn_names.append(k)
            bn_norms.append(float(v.double().norm()))
    rec["bn_names"] = np.array(bn_names)
    rec["bn_norms"] = np.array(bn_norms)
    np.savez_compressed(OUT, **rec)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; loss", float(total), "items", items.numpy(), "params", len(names))


if __name__ == "__main__":
    main()
