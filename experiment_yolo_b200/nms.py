"""Device-side post-processing: the reference's `non_max_suppression` as ONE CUDA launch (csrc/ldconv_nms.cu).

Mirrors /root/reference/ultralytics/utils/ops.py:292-427 for the path the predictor / validator take (single label, no masks,
not rotated, classes=None, labels=()), including the fork's own `soft_nms` (ops.py:260-290) with its quirks -- see the kernel's
header.  Same argument names and meaning; unsupported options raise NotImplementedError, a CPU tensor raises RuntimeError
(no CPU fallback).  `nms_padded` is the graph-capturable form the pipelined predictor uses: fixed-shape outputs, no host sync.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _lib

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def nms_workspace(B: int, A: int, max_nms: int, device) -> torch.Tensor:
    n = int(_lib.load().ldconv_nms_workspace_bytes(B, min(A, max_nms)))
    return torch.empty(max(n, 16), device=device, dtype=torch.uint8)


def nms_padded(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45, agnostic: bool = False,
               max_det: int = 300, nc: int = 0, max_nms: int = 30000, max_wh: float = 7680.0,
               out: Optional[torch.Tensor] = None, count: Optional[torch.Tensor] = None,
               workspace: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """prediction (B, 4+nc, A) bf16 / fp32 on the GPU -> out (B, max_det, 6) fp32 rows (x1, y1, x2, y2, conf, cls) in keep
    order and count (B,) int32 (negative: more than max_nms candidates, raise conf_thres).  Asynchronous, no host sync."""
    if not prediction.is_cuda:
        raise RuntimeError("experiment_yolo_b200.nms needs a CUDA tensor; there is no CPU fallback")
    if prediction.dim() != 3 or prediction.dtype not in _DT:
        raise TypeError(f"nms: expected a (B, 4+nc, A) float32 / bfloat16 tensor, got {tuple(prediction.shape)} {prediction.dtype}")
    if not 0 <= conf_thres <= 1 or not 0 <= iou_thres <= 1:
        raise AssertionError(f"Invalid thresholds conf={conf_thres} iou={iou_thres}, valid values are between 0.0 and 1.0")
    B, ch, A = prediction.shape
    nc = nc or (ch - 4)
    if ch - nc - 4 != 0:
        raise NotImplementedError("nms: mask channels are not supported")
    pred = prediction.contiguous()
    dev = pred.device
    if out is None:
        out = torch.zeros((B, max_det, 6), device=dev, dtype=torch.float32)
    if count is None:
        count = torch.zeros((B,), device=dev, dtype=torch.int32)
    if workspace is None:
        workspace = nms_workspace(B, A, max_nms, dev)
    _lib.check(_lib.load().ldconv_nms(pred.data_ptr(), out.data_ptr(), count.data_ptr(), workspace.data_ptr(), workspace.numel(), B, A,
                                      nc, float(conf_thres), float(iou_thres), int(bool(agnostic)), int(max_det), int(max_nms),
                                      float(max_wh), _DT[pred.dtype], torch.cuda.current_stream().cuda_stream), "ldconv_nms")
    return out, count


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False, labels=(),
                        max_det=300, nc=0, max_time_img=0.05, max_nms=30000, max_wh=7680, rotated=False) -> List[torch.Tensor]:
    """Drop-in for ultralytics.utils.ops.non_max_suppression (ops.py:292): list of (n_i, 6) tensors per image."""
    if isinstance(prediction, (list, tuple)):      # (inference_out, loss_out) in validation (ops.py:340-341)
        prediction = prediction[0]
    if classes is not None or multi_label or (labels and any(len(l) for l in labels)) or rotated:
        raise NotImplementedError("nms: classes / multi_label / labels / rotated are not covered by the CUDA kernel")
    out, count = nms_padded(prediction, conf_thres, iou_thres, agnostic, max_det, nc, max_nms, float(max_wh))
    cnt = count.tolist()      # the one host sync of the eager API
    if any(c < 0 for c in cnt):
        raise RuntimeError(f"nms: an image has {-min(cnt)} candidates above conf_thres={conf_thres}, more than max_nms={max_nms}; "
                           "the confidence-sorted truncation of ops.py:395-396 is not implemented -- raise conf_thres")
    return [out[b, :c] for b, c in enumerate(cnt)]
