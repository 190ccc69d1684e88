"""numpy restatement of the reference's post-processing (TEST INFRASTRUCTURE ONLY -- never imported by the product).

Follows /root/reference/ultralytics/utils/ops.py:292-427 (`non_max_suppression`: single label, no masks, not rotated,
classes=None, labels=()) and the fork's `soft_nms` (ops.py:260-290) that line 407 calls, with its quirks: candidates in
anchor order, in-place score decay exp(-IoU^2 / 0.5) for IoU > iou_thres, survivors need score > 0.25, arg-max swapped to
the front, the last survivor is never kept (`while order.numel() > 1`), no decay when only two boxes are left (0-d squeeze).
Pinned by tests/golden/nms_*.npz (minted from the reference by oracle/gen_nms_golden.py; tests/test_nms_cpu.py).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def _iou_one_to_many(b1, b2, eps=F(1e-7)):
    """bbox_iou_for_nms (ops.py:188-202), xyxy, fp32"""
    w1, h1 = b1[2] - b1[0], b1[3] - b1[1] + eps
    w2, h2 = b2[:, 2] - b2[:, 0], b2[:, 3] - b2[:, 1] + eps
    iw = np.clip(np.minimum(b1[2], b2[:, 2]) - np.maximum(b1[0], b2[:, 0]), 0, None)
    ih = np.clip(np.minimum(b1[3], b2[:, 3]) - np.maximum(b1[1], b2[:, 1]), 0, None)
    inter = (iw * ih).astype(F)
    union = (w1 * h1 + w2 * h2 - inter + eps).astype(F)
    return (inter / union).astype(F)


def soft_nms(boxes, scores, iou_thresh, sigma=F(0.5), score_threshold=F(0.25)):
    """ops.py:260-290; `scores` is modified in place like the reference's view of x[:, 4]"""
    order = np.arange(scores.shape[0])
    keep = []
    while order.size > 1:
        i = order[0]
        keep.append(int(i))
        rest = order[1:]
        if rest.size > 1:      # a single remaining box makes `iou` 0-d and the nonzero().squeeze() empty: no decay
            iou = _iou_one_to_many(boxes[i], boxes[rest])
            hit = np.nonzero(iou > iou_thresh)[0]
            if hit.size:
                scores[rest[hit]] = (scores[rest[hit]] * np.exp(-(iou[hit] * iou[hit]) / sigma).astype(F)).astype(F)
        alive = np.nonzero(scores[rest] > score_threshold)[0]
        if alive.size == 0:
            break
        best = int(np.argmax(scores[rest[alive]]))
        if best != 0:
            alive[[0, best]] = alive[[best, 0]]
        order = rest[alive]
    return keep


def non_max_suppression(pred, conf_thres=0.25, iou_thres=0.45, agnostic=False, max_det=300, max_nms=30000, max_wh=7680):
    """pred (B, 4+nc, A) fp32 -> list of (n_i, 6) fp32 arrays (x1, y1, x2, y2, conf, cls)"""
    pred = np.asarray(pred, dtype=F)
    B, ch, A = pred.shape
    out = []
    for b in range(B):
        p = pred[b].T                                  # (A, 4+nc)
        cls = p[:, 4:]
        conf = cls.max(1)
        j = cls.argmax(1)
        sel = conf > F(conf_thres)
        xy, wh = p[sel, :2], p[sel, 2:4]
        half = (wh / F(2)).astype(F)
        box = np.concatenate([xy - half, xy + half], 1).astype(F)      # xywh2xyxy
        conf, j = conf[sel].copy(), j[sel]
        if box.shape[0] == 0:
            out.append(np.zeros((0, 6), F))
            continue
        if box.shape[0] > max_nms:
            raise NotImplementedError("more than max_nms candidates")
        c = (j.astype(F) * F(0 if agnostic else max_wh))[:, None]
        keep = soft_nms((box + c).astype(F), conf, F(iou_thres))[:max_det]
        out.append(np.concatenate([box[keep], conf[keep, None], j[keep, None].astype(F)], 1).astype(F).reshape(-1, 6))
    return out
