"""Mint NMS golden vectors FROM THE REFERENCE's own `non_max_suppression` (authoring container only).

Imports the unmodified /root/reference/ultralytics/utils/ops.py (stub importer of oracle/gen_model_golden.py for the absent
third-party roots), runs `non_max_suppression` (ops.py:292-427, which calls the fork's `soft_nms`, ops.py:260-290) on seeded
synthetic head outputs (B, 4+nc, anchors) and stores inputs + per-image outputs under tests/golden/nms_*.npz.

    python oracle/gen_nms_golden.py

TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle.gen_model_golden import load_reference_tasks  # noqa: E402

OUT_DIR = os.path.join(ROOT, "tests", "golden")


def synth(seed, B, A, nc, objects, hot_frac, img=640.0, dtype=torch.float32):
    """head-like predictions: boxes jittered around `objects` centres (so that many overlap), a fraction `hot_frac` of the
    anchors carries one confident class"""
    g = torch.Generator().manual_seed(seed)
    cxy = torch.rand(B, objects, 2, generator=g) * (img - 80) + 40
    wh = torch.rand(B, objects, 2, generator=g) * 60 + 12
    ocls = torch.randint(0, nc, (B, objects), generator=g)
    which = torch.randint(0, objects, (B, A), generator=g)
    bi = torch.arange(B)[:, None]
    boxes = torch.cat([cxy[bi, which] + torch.randn(B, A, 2, generator=g) * 4.0,
                       wh[bi, which] * (1 + 0.15 * torch.randn(B, A, 2, generator=g)).clamp(0.5, 1.5)], -1)      # (B,A,4) xywh
    scores = torch.rand(B, A, nc, generator=g) * 0.2
    hot = torch.rand(B, A, generator=g) < hot_frac
    conf = torch.rand(B, A, generator=g) * 0.7 + 0.27
    cls_of = ocls[bi, which]
    scores[bi.expand(B, A)[hot], torch.arange(A).expand(B, A)[hot], cls_of[hot]] = conf[hot]
    pred = torch.cat([boxes, scores], -1).transpose(1, 2).contiguous()      # (B, 4+nc, A)
    return pred.to(dtype).float()      # values exactly representable in `dtype`


def main():
    warnings.filterwarnings("ignore")
    load_reference_tasks()
    from ultralytics.utils import ops as R
    cases = {
        "default": dict(pred=synth(1, 3, 2100, 6, 30, 0.12), kw=dict(conf_thres=0.25, iou_thres=0.45)),
        "bf16_agnostic": dict(pred=synth(2, 2, 2100, 6, 12, 0.2, dtype=torch.bfloat16), kw=dict(conf_thres=0.3, iou_thres=0.5, agnostic=True)),
        "many_small_maxdet": dict(pred=synth(3, 2, 4200, 6, 60, 0.3), kw=dict(conf_thres=0.25, iou_thres=0.45, max_det=20)),
        "lowconf": dict(pred=synth(4, 1, 1000, 6, 10, 0.05), kw=dict(conf_thres=0.05, iou_thres=0.6)),
    }
    # edge images: no candidate, exactly one candidate (the reference returns nothing), exactly two
    edge = synth(5, 3, 300, 6, 5, 0.0)
    edge[1, 4 + 2, 17] = 0.9
    edge[2, 4 + 1, 40] = 0.8
    edge[2, 4 + 1, 41] = 0.7
    edge[2, :4, 41] = edge[2, :4, 40] + 1.0
    cases["edge_counts"] = dict(pred=edge, kw=dict(conf_thres=0.25, iou_thres=0.45))
    for name, c in cases.items():
        pred = c["pred"]
        out = R.non_max_suppression(pred.clone(), **c["kw"])
        counts = np.array([o.shape[0] for o in out], dtype=np.int32)
        kw = c["kw"]
        np.savez_compressed(os.path.join(OUT_DIR, f"nms_{name}.npz"), pred=pred.numpy(), counts=counts,
                            **{f"out{i}": o.numpy() for i, o in enumerate(out)},
                            conf_thres=np.float32(kw.get("conf_thres", 0.25)), iou_thres=np.float32(kw.get("iou_thres", 0.45)),
                            agnostic=np.int32(kw.get("agnostic", False)), max_det=np.int32(kw.get("max_det", 300)))
        cand = (pred[:, 4:].amax(1) > kw["conf_thres"]).sum(1).tolist()
        print(name, "candidates", cand, "kept", counts.tolist())


if __name__ == "__main__":
    main()
