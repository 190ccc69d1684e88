"""DEAL-YOLO detection loss (BASELINE.json config 4): task-aligned assignment + BCE + Wise-IoU v3 + NWD + DFL.

Host-side restatement, written from scratch, of the reference training criterion for the yolov8-LD-P2 head:
  * `v8DetectionLoss.__call__/compute_loss/preprocess/bbox_decode`  /root/reference/ultralytics/utils/loss.py:296-433
  * `BboxLoss.forward/_df_loss` with `use_wiseiou`, `nwd_loss`, `iou_ratio`  utils/loss.py:187-250
  * `TaskAlignedAssigner` (topk 10, alpha 0.5, beta 6)  utils/tal.py:13-290, `make_anchors/dist2bbox/bbox2dist`  :294-324
  * `bbox_iou(CIoU=True)`  utils/metrics.py:75-128, `wasserstein_loss`  :540-565, `WiseIouLoss('WIoU', monotonous=False)`  :567-645
It is plumbing around the LDConv hot path, not a kernel: plain PyTorch, any device.

Same arithmetic, different data flow.  The reference indexes with boolean masks (`pred_bboxes[fg_mask]`, `overlaps[mask_gt]`,
`if fg_mask.sum():`, python `max(target_scores.sum(), 1)`), each of which is a device->host synchronisation and a
dynamically shaped tensor; on a B200 those stalls cost more than the arithmetic.  Here every tensor has a static shape:
  * the assigner works on dense (b, n_gt, n_anchors) tensors with `torch.where` instead of masked assignment, and always
    resolves multi-assigned anchors (a no-op when there are none) instead of branching on `fg_mask.max() > 1`;
  * the box terms run on a fixed-capacity compaction of the foreground anchors: an image has at most n_gt * topk positives,
    so `topk` of the 0/1 foreground mask yields their indices without `nonzero()`; padding rows carry weight 0 and harmless
    unit boxes (so no NaN can leak into the gradients through a masked branch);
  * `iou_mean` (Wise-IoU's running mean) is updated with `torch.where(count > 0, ...)`.
The only host synchronisation left is the optional `n_max = counts.max()` of the target padding; pass `max_boxes` to avoid it.
tests/test_loss_cpu.py pins every output and gradient against fixtures minted from the reference (oracle/gen_loss_golden.py).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


def make_anchors(shapes, strides, offset=0.5, device=None, dtype=torch.float32):
    """Anchor centres (in grid units) and per-anchor stride for feature maps of `shapes` [(h, w), ...]  (tal.py:294-306)."""
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sx = torch.arange(w, device=device, dtype=dtype) + offset
        sy = torch.arange(h, device=device, dtype=dtype) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s), device=device, dtype=dtype))
    return torch.cat(pts), torch.cat(st)


def ciou(b1, b2, eps=1e-7):
    """Complete IoU of xyxy boxes, broadcasting over leading dims; returns (..., 1)  (metrics.py:103-128)."""
    b1_x1, b1_y1, b1_x2, b1_y2 = b1.chunk(4, -1)
    b2_x1, b2_y1, b2_x2, b2_y2 = b2.chunk(4, -1)
    w1, h1 = b1_x2 - b1_x1, b1_y2 - b1_y1 + eps
    w2, h2 = b2_x2 - b2_x1, b2_y2 - b2_y1 + eps
    inter = (torch.minimum(b1_x2, b2_x2) - torch.maximum(b1_x1, b2_x1)).clamp(min=0) * \
            (torch.minimum(b1_y2, b2_y2) - torch.maximum(b1_y1, b2_y1)).clamp(min=0)
    union = w1 * h1 + w2 * h2 - inter + eps
    iou = inter / union
    cw = torch.maximum(b1_x2, b2_x2) - torch.minimum(b1_x1, b2_x1)
    ch = torch.maximum(b1_y2, b2_y2) - torch.minimum(b1_y1, b2_y1)
    c2 = cw ** 2 + ch ** 2 + eps
    rho2 = ((b2_x1 + b2_x2 - b1_x1 - b1_x2) ** 2 + (b2_y1 + b2_y2 - b1_y1 - b1_y2) ** 2) / 4
    v = (4 / math.pi ** 2) * (torch.atan(w2 / h2) - torch.atan(w1 / h1)).pow(2)
    with torch.no_grad():
        alpha = v / (v - iou + (1 + eps))
    return iou - (rho2 / c2 + v * alpha)


class TaskAlignedAssigner:
    """tal.py:13-290 on dense, statically shaped tensors (no boolean-mask indexing, no data-dependent branches)."""

    def __init__(self, topk=10, num_classes=80, alpha=0.5, beta=6.0, eps=1e-9, fused=True):
        self.topk, self.nc, self.alpha, self.beta, self.eps = topk, num_classes, alpha, beta, eps
        self.fused = fused      # CUDA tensors: the three kernels of csrc/ldconv_tal.cu instead of ~60 dense PyTorch passes

    def _fused(self, pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, valid_gt):
        """Same assignment through libldconv_b200 (ldconv_tal_metric / ldconv_tal_assign); torch.topk stays a library call."""
        from . import _lib
        L = _lib.load()
        b, na, nc = pd_scores.shape
        n = gt_bboxes.shape[1]
        dev = pd_scores.device
        st = torch.cuda.current_stream(dev).cuda_stream
        f32 = dict(device=dev, dtype=torch.float32)
        scores, boxes = pd_scores.float().contiguous(), pd_bboxes.float().contiguous()
        anc, gtb = anc_points.float().contiguous(), gt_bboxes.float().contiguous()
        labels = gt_labels.reshape(b, n).to(torch.int32).contiguous()
        valid = valid_gt.reshape(b, n).to(torch.uint8).contiguous()
        align, overlaps = torch.empty((b, n, na), **f32), torch.empty((b, n, na), **f32)
        _lib.check(L.ldconv_tal_metric(scores.data_ptr(), boxes.data_ptr(), anc.data_ptr(), labels.data_ptr(), gtb.data_ptr(),
                                       valid.data_ptr(), align.data_ptr(), overlaps.data_ptr(), b, na, n, nc, self.alpha, self.beta,
                                       self.eps, st), "ldconv_tal_metric")
        k = min(self.topk, na)
        idx = torch.topk(align, k, dim=-1).indices.contiguous()
        mask_ws = torch.empty((b, n, na), device=dev, dtype=torch.uint8)
        fg = torch.empty((b, na), device=dev, dtype=torch.uint8)
        gt_idx = torch.empty((b, na), device=dev, dtype=torch.int64)
        align_sel = torch.empty((b, na), **f32)
        pos_align, pos_over = torch.empty((b, n), **f32), torch.empty((b, n), **f32)
        _lib.check(L.ldconv_tal_assign(idx.data_ptr(), anc.data_ptr(), gtb.data_ptr(), valid.data_ptr(), align.data_ptr(),
                                       overlaps.data_ptr(), mask_ws.data_ptr(), fg.data_ptr(), gt_idx.data_ptr(), align_sel.data_ptr(),
                                       pos_align.data_ptr(), pos_over.data_ptr(), b, na, n, k, self.eps, st), "ldconv_tal_assign")
        fgb = fg.bool()
        flat = gt_idx + torch.arange(b, device=dev).view(-1, 1) * n
        target_labels = gt_labels.long().flatten()[flat].clamp(min=0)
        target_bboxes = gt_bboxes.reshape(-1, 4)[flat]
        norm = align_sel * pos_over.flatten()[flat] / (pos_align.flatten()[flat] + self.eps)      # 0 on the background
        target_scores = F.one_hot(target_labels, self.nc).to(align.dtype) * norm.unsqueeze(-1)
        return target_labels, target_bboxes, target_scores, fgb, gt_idx

    @torch.no_grad()
    def __call__(self, pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, mask_gt):
        """pd_scores (b,na,nc) in [0,1], pd_bboxes (b,na,4) xyxy px, anc_points (na,2) px, gt_labels (b,n,1), gt_bboxes
        (b,n,4) xyxy px, mask_gt (b,n,1) -> target_labels (b,na), target_bboxes (b,na,4), target_scores (b,na,nc),
        fg_mask (b,na) bool, target_gt_idx (b,na)."""
        b, na, _ = pd_scores.shape
        n = gt_bboxes.shape[1]
        if n == 0:                                                                        # tal.py:61-69
            return (torch.full((b, na), self.nc, device=pd_scores.device, dtype=pd_scores.dtype), torch.zeros_like(pd_bboxes),
                    torch.zeros_like(pd_scores), torch.zeros((b, na), device=pd_scores.device, dtype=torch.bool),
                    torch.zeros((b, na), device=pd_scores.device, dtype=torch.long))
        valid_gt = mask_gt.bool()                                                         # (b,n,1)
        if pd_scores.is_cuda and self.fused and n <= 256:
            return self._fused(pd_scores, pd_bboxes, anc_points, gt_labels, gt_bboxes, valid_gt)
        # anchors strictly inside each gt box (tal.py:226-243)
        lt, rb = gt_bboxes.view(b, n, 1, 4).chunk(2, -1)
        ap = anc_points.view(1, 1, na, 2)
        in_gts = torch.minimum((ap - lt).amin(-1), (rb - ap).amin(-1)) > self.eps         # (b,n,na)
        cand = in_gts & valid_gt
        # alignment metric = score^alpha * CIoU^beta on the candidates, 0 elsewhere (tal.py:98-122)
        labels = gt_labels.long().clamp(0, self.nc - 1).view(b, n, 1)
        scores = torch.gather(pd_scores.transpose(1, 2), 1, labels.expand(b, n, na))      # (b,n,na): score of the gt's class
        zero = scores.new_zeros(())
        overlaps = torch.where(cand, ciou(gt_bboxes.view(b, n, 1, 4), pd_bboxes.view(b, 1, na, 4)).squeeze(-1).clamp(min=0), zero)
        align = torch.where(cand, scores, zero).pow(self.alpha) * overlaps.pow(self.beta)
        # top-k anchors per gt (tal.py:124-157); padded gts point all k indices at anchor 0, which the count > 1 filter drops
        k = min(self.topk, na)
        idx = torch.topk(align, k, dim=-1).indices
        idx = torch.where(valid_gt.expand(-1, -1, k), idx, torch.zeros_like(idx))
        count = torch.zeros((b, n, na), device=align.device, dtype=torch.int32)
        count.scatter_add_(-1, idx, torch.ones_like(idx, dtype=torch.int32))
        mask_pos = (count == 1) & cand                                                    # (b,n,na)
        # an anchor claimed by several gts goes to the one with the highest overlap (tal.py:245-272)
        mp = mask_pos.to(align.dtype)
        multi = (mp.sum(-2, keepdim=True) > 1).expand(-1, n, -1)
        is_max = torch.zeros_like(mp).scatter_(1, overlaps.argmax(1, keepdim=True), 1.0)
        mp = torch.where(multi, is_max, mp)
        fg = mp.sum(-2)                                                                   # (b,na) in {0,1}
        gt_idx = mp.argmax(-2)                                                            # (b,na)
        # targets of the assigned gt (tal.py:159-208)
        flat = gt_idx + torch.arange(b, device=gt_idx.device).view(-1, 1) * n
        target_labels = gt_labels.long().flatten()[flat].clamp(min=0)
        target_bboxes = gt_bboxes.reshape(-1, 4)[flat]
        fgb = fg > 0
        target_scores = F.one_hot(target_labels, self.nc).to(align.dtype) * fgb.unsqueeze(-1)
        # normalise by the best alignment / overlap of each gt (tal.py:83-88)
        align = align * mp
        pos_align = align.amax(-1, keepdim=True)
        pos_over = (overlaps * mp).amax(-1, keepdim=True)
        norm = (align * pos_over / (pos_align + self.eps)).amax(-2).unsqueeze(-1)
        return target_labels, target_bboxes, target_scores * norm, fgb, gt_idx


class WiseIoU(nn.Module):
    """Wise-IoU v3 ('WIoU', monotonous=False) with its running mean of the IoU loss (metrics.py:567-645)."""
    momentum, alpha, delta = 1e-2, 1.7, 2.7

    def __init__(self):
        super().__init__()
        self.register_buffer("iou_mean", torch.tensor(1.0))

    def forward(self, pred, target, valid):
        """pred, target (m,4) xyxy; valid (m,) bool marks real rows (padding rows must hold finite boxes) -> (m,)."""
        pwh, twh = pred[:, 2:] - pred[:, :2], target[:, 2:] - target[:, :2]
        mn, mx = torch.minimum(pred, target), torch.maximum(pred, target)
        s_inter = torch.relu(mn[:, 2:] - mx[:, :2]).prod(-1)
        s_union = pwh.prod(-1) + twh.prod(-1) - s_inter
        wh_box = mx[:, 2:] - mn[:, :2]
        l2_box = wh_box.square().sum(-1)
        l2_center = ((pred[:, :2] + pred[:, 2:]) / 2 - (target[:, :2] + target[:, 2:]) / 2).square().sum(-1)
        iou = 1 - s_inter / s_union                                                       # the IoU *loss*
        if self.training:                                                                 # metrics.py:621-623
            cnt = valid.sum()
            mean = (iou.detach() * valid).sum() / cnt.clamp(min=1)
            new = self.iou_mean * (1 - self.momentum) + self.momentum * mean
            self.iou_mean.copy_(torch.where(cnt > 0, new, self.iou_mean))
        loss = torch.exp(l2_center / l2_box.detach()) * iou                               # _WIoU, metrics.py:643-645
        beta = iou.detach() / self.iou_mean                                               # _scaled_loss, :629-638
        return loss * (beta / (self.delta * torch.pow(self.alpha, beta - self.delta)))


def nwd(pred, target, eps=1e-7, constant=12.8):
    """Normalised Gaussian Wasserstein distance of xyxy boxes, (m,4) -> (m,)  (metrics.py:540-565)."""
    w1, h1 = pred[:, 2] - pred[:, 0], pred[:, 3] - pred[:, 1] + eps
    w2, h2 = target[:, 2] - target[:, 0], target[:, 3] - target[:, 1] + eps
    cx1, cy1, cx2, cy2 = pred[:, 0] + w1 / 2, pred[:, 1] + h1 / 2, target[:, 0] + w2 / 2, target[:, 1] + h2 / 2
    center = (cx1 - cx2) ** 2 + (cy1 - cy2) ** 2 + eps
    wh = ((w1 - w2) ** 2 + (h1 - h2) ** 2) / 4
    return torch.exp(-torch.sqrt(center + wh) / constant)


class DealYoloLoss(nn.Module):
    """`v8DetectionLoss` with `bbox_loss.use_wiseiou = True`, `nwd_loss = True`, `iou_ratio = 0.5`.

    __call__(feats, batch) -> (loss.sum() * batch_size, loss.detach()) like the reference (loss.py:356-361): feats = the three
    raw head maps (b, 4*reg_max + nc, h, w); batch = {'batch_idx' (t,), 'cls' (t,) or (t,1), 'bboxes' (t,4) xywh normalised}.
    """

    def __init__(self, nc=6, reg_max=16, strides=(4.0, 8.0, 16.0), box=7.5, cls=0.5, dfl=1.5, topk=10, use_wiseiou=True,
                 nwd_loss=True, iou_ratio=0.5, max_boxes=None, fused_assigner=True):
        super().__init__()
        self.nc, self.reg_max, self.no = nc, reg_max, nc + 4 * reg_max
        self.strides = [float(s) for s in strides]
        self.hyp = SimpleNamespace(box=box, cls=cls, dfl=dfl)
        self.assigner = TaskAlignedAssigner(topk=topk, num_classes=nc, alpha=0.5, beta=6.0, fused=fused_assigner)
        self.topk = topk
        self.use_wiseiou, self.nwd_loss, self.iou_ratio = use_wiseiou, nwd_loss, iou_ratio
        self.wiou_loss = WiseIoU()
        self.max_boxes = max_boxes
        self.sparse_head = True      # CUDA: dense no-grad decode kernel + foreground-row decode (see forward); False = the dense torch path
        self.register_buffer("proj", torch.arange(reg_max, dtype=torch.float32), persistent=False)

    # ---- targets: (t, 6) rows -> (b, n_max, 5) padded [cls, xyxy px]  (loss.py:329-345) ---------------------------------
    def preprocess(self, batch, b, imgsz_hw, device):
        idx = batch["batch_idx"].to(device).view(-1).long()
        t = idx.numel()
        if t == 0:
            return torch.zeros((b, 0, 5), device=device)
        rows = torch.cat((batch["cls"].to(device).view(-1, 1).float(), batch["bboxes"].to(device).view(-1, 4).float()), 1)
        counts = torch.bincount(idx, minlength=b)
        n_max = self.max_boxes if self.max_boxes is not None else int(counts.max())
        order = torch.sort(idx, stable=True).indices
        start = torch.cumsum(counts, 0) - counts
        pos = torch.arange(t, device=device) - start[idx[order]]
        out = torch.zeros((b, n_max, 5), device=device)
        if self.max_boxes is not None:
            # a fixed `max_boxes` avoids the host sync of counts.max(); targets beyond it are dropped by a mask (an unchecked
            # `out[idx, pos] = ...` with pos >= n_max is a device-side assert that poisons the CUDA context)
            keep = pos < n_max
            out[idx[order][keep], pos[keep]] = rows[order][keep]
        else:
            out[idx[order], pos] = rows[order]
        h, w = imgsz_hw
        xywh = out[..., 1:5] * torch.tensor([w, h, w, h], device=device, dtype=out.dtype)
        half = xywh[..., 2:] / 2
        out[..., 1:5] = torch.cat((xywh[..., :2] - half, xywh[..., :2] + half), -1)
        return out

    def forward(self, feats, batch):
        feats = list(feats[:len(self.strides)])
        b = feats[0].shape[0]
        dev = feats[0].device
        shapes = [tuple(f.shape[2:]) for f in feats]
        imgsz = (shapes[0][0] * self.strides[0], shapes[0][1] * self.strides[0])
        anchor_points, stride_tensor = make_anchors(shapes, self.strides, 0.5, device=dev)
        targets = self.preprocess(batch, b, imgsz, dev)
        gt_labels, gt_bboxes = targets.split((1, 4), 2)
        mask_gt = gt_bboxes.sum(2, keepdim=True) > 0
        sparse = self.sparse_head and feats[0].is_cuda and self.reg_max == 16 and feats[0].dtype in (torch.bfloat16, torch.float32)
        if sparse:
            # CUDA path: ONE anchor-major concatenation of the maps in their own dtype (zero-copy NHWC views of channels_last maps in,
            # (b, na, no) out).  The assigner's inputs -- decoded boxes and sigmoid scores of ALL anchors, which carry no gradient --
            # come from one kernel (ldconv_head_decode_rows); the criterion differentiates through the foreground rows only, so the
            # dense fp32 copies, softmax, matmul and their backward over 33600 anchors per image disappear (~14 -> ~7 ms at batch 128).
            from . import _lib
            x = torch.cat([f.permute(0, 2, 3, 1).reshape(b, -1, self.no) for f in feats], 1)          # (b, na, no)
            na = x.shape[1]
            pred_scores = x[..., self.reg_max * 4:].float()                                          # (b, na, nc), differentiable
            boxes_all = torch.empty((b, na, 4), device=dev, dtype=torch.float32)
            scores_sig = torch.empty((b, na, self.nc), device=dev, dtype=torch.float32)
            xd = x.detach()
            _lib.check(_lib.load().ldconv_head_decode_rows(
                xd.data_ptr(), anchor_points.contiguous().data_ptr(), boxes_all.data_ptr(), scores_sig.data_ptr(), b * na, na, self.nc,
                self.reg_max, _lib.BF16 if x.dtype == torch.bfloat16 else _lib.F32, torch.cuda.current_stream(dev).cuda_stream),
                "ldconv_head_decode_rows")
            pred_distri = pred_bboxes = None
        else:
            # one concatenation in the maps' own dtype (bf16 under autocast), one fp32 conversion fused with the transposition
            x = torch.cat([f.reshape(b, self.no, -1) for f in feats], 2)
            pred_distri, pred_scores = x.split((self.reg_max * 4, self.nc), 1)
            pred_scores = pred_scores.permute(0, 2, 1).float().contiguous()                    # (b,na,nc)
            pred_distri = pred_distri.permute(0, 2, 1).float().contiguous()                    # (b,na,4*reg_max)
            na = pred_scores.shape[1]
            # decode: DFL expectation -> ltrb -> xyxy in grid units (loss.py:347-354, tal.py:309-318)
            dist = pred_distri.view(b, na, 4, self.reg_max).softmax(3).matmul(self.proj)
            pred_bboxes = torch.cat((anchor_points - dist[..., :2], anchor_points + dist[..., 2:]), -1)
            boxes_all, scores_sig = pred_bboxes.detach(), pred_scores.detach().sigmoid()

        _, target_bboxes, target_scores, fg_mask, _ = self.assigner(
            scores_sig, boxes_all * stride_tensor, anchor_points * stride_tensor, gt_labels, gt_bboxes, mask_gt.to(gt_bboxes.dtype))
        tss = target_scores.sum().clamp(min=1)

        loss = torch.zeros(3, device=dev)
        loss_cls = F.binary_cross_entropy_with_logits(pred_scores, target_scores, reduction="none").sum() / tss

        # ---- box terms on the compacted foreground anchors (at most n_gt * topk per image) ---------------------------------
        n_gt = gt_bboxes.shape[1]
        cap = min(na, max(1, n_gt * self.topk))
        sel_v, sel = torch.topk(fg_mask.to(torch.float32), cap, dim=1)                     # the 1s come first
        valid = (sel_v > 0).reshape(-1)                                                    # (b*cap,)
        take = lambda t: torch.gather(t, 1, sel.unsqueeze(-1).expand(-1, -1, t.shape[-1])).reshape(b * cap, t.shape[-1])
        unit = torch.tensor([0.0, 0.0, 1.0, 1.0], device=dev)
        vcol = valid.unsqueeze(-1)
        ap = anchor_points.unsqueeze(0).expand(b, -1, -1)
        apc = take(ap)
        if sparse:      # decode the foreground rows only (same arithmetic as the dense decode, row by row)
            pd_rows = take(x[..., : self.reg_max * 4]).float()                                       # (b*cap, 4*reg_max)
            dist = pd_rows.view(-1, 4, self.reg_max).softmax(-1).matmul(self.proj)
            pb_rows = torch.cat((apc - dist[:, :2], apc + dist[:, 2:]), -1)
        else:
            pd_rows, pb_rows = take(pred_distri), take(pred_bboxes)
        pb = torch.where(vcol, pb_rows, unit)
        tb = torch.where(vcol, take(target_bboxes / stride_tensor), unit)
        weight = torch.where(valid, take(target_scores).sum(-1), torch.zeros((), device=dev))
        if self.use_wiseiou:
            l_iou = self.wiou_loss(pb, tb, valid)
        else:
            l_iou = 1.0 - ciou(pb, tb).squeeze(-1)
        loss_iou = (l_iou * weight).sum() / tss
        if self.nwd_loss:
            loss_nwd = ((1.0 - nwd(pb, tb)) * weight).sum() / tss
            loss_iou = self.iou_ratio * loss_iou + (1 - self.iou_ratio) * loss_nwd
        # DFL (loss.py:226-250, tal.py:321-324)
        ltrb = torch.cat((apc - tb[:, :2], tb[:, 2:] - apc), -1).clamp(0, self.reg_max - 1 - 0.01)
        ltrb = torch.where(vcol, ltrb, torch.zeros((), device=dev))
        logp = F.log_softmax(pd_rows.view(-1, 4, self.reg_max), -1)
        tl = ltrb.long()
        wl = (tl + 1).to(ltrb.dtype) - ltrb
        ce_l = -logp.gather(-1, tl.unsqueeze(-1)).squeeze(-1)
        ce_r = -logp.gather(-1, (tl + 1).unsqueeze(-1)).squeeze(-1)
        loss_dfl = ((ce_l * wl + ce_r * (1 - wl)).mean(-1) * weight).sum() / tss

        loss = torch.stack((loss_iou * self.hyp.box, loss_cls * self.hyp.cls, loss_dfl * self.hyp.dfl))
        return loss.sum() * b, loss.detach()


def synthetic_uav_targets(b, boxes_per_image=16, nc=6, seed=0, device="cpu"):
    """Synthetic UAV-shaped targets of SURVEY.md 8d config 4: per image 16 boxes, cls ~ U{0..nc-1}, centre ~ U(0.05, 0.95)^2,
    w, h ~ U(0.01, 0.05) of the image; batch dict keys as data/dataset.py:207-223."""
    g = torch.Generator().manual_seed(seed)
    t = b * boxes_per_image
    batch_idx = torch.arange(b).repeat_interleave(boxes_per_image).float()
    cls = torch.randint(0, nc, (t, 1), generator=g).float()
    xy = torch.rand((t, 2), generator=g) * 0.9 + 0.05
    wh = torch.rand((t, 2), generator=g) * 0.04 + 0.01
    return {"batch_idx": batch_idx.to(device), "cls": cls.to(device), "bboxes": torch.cat((xy, wh), 1).to(device)}
