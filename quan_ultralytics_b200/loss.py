"""`v8OBBLoss` (ultralytics/utils/loss.py:853-1050) with static shapes and no host synchronisation, so that the loss — 9-14 ms of
host-bound Python per QUAN-YOLO11n step in the reference (boolean-mask indexing, `.item()`, CPU-built index tensors) — can be
captured into the training step's CUDA graph (graphs.GraphedTrainStep(capture_loss=True)).  SURVEY §8(f) rank 4.

Same interface as the reference criterion: `loss_fn = OBBLossStatic(model); total, items = loss_fn(preds, batch)` with
`preds = (feats, pred_angle)` from the OBB head in training mode, returning `(loss.sum() * batch_size, loss.detach())` with
items = (box, cls, dfl, quaternion_angle).  Differences, all forced by static shapes:
  * targets come padded: batch["targets"] [B, n_max, 6] = (cls, x, y, w, h, theta) normalised + batch["target_mask"] [B, n_max]
    (`pad_targets` builds them on the host from the reference's batch_idx / cls / bboxes lists, loss.py:961-968 `preprocess`);
  * the assigner is `quan_rotated_tal_assign` (csrc/tal.cu); everything selected with `tensor[fg_mask]` in the reference is a
    masked dense sum here (same value up to summation order);
  * the differentiable tail (BCE, ProbIoU, DFL, quaternion angular loss: loss.py:995-1033, :364-378, :306-329) is dense torch
    arithmetic under autograd, in fp32.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import check


def pad_targets(batch: Dict[str, torch.Tensor], batch_size: int, n_max: int | None = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """loss.py:925-939 `preprocess` without the per-image device loop: (targets [B, n_max, 6], mask [B, n_max]) from the
    reference's flat lists; boxes stay normalised (the loss scales them).  Host-side (run it where the lists live)."""
    idx = batch["batch_idx"].view(-1).long().cpu()
    cls = batch["cls"].view(-1, 1).float().cpu()
    box = batch["bboxes"].view(-1, 5).float().cpu()
    counts = torch.bincount(idx, minlength=batch_size)
    n_max = int(counts.max()) if n_max is None else n_max
    out = torch.zeros(batch_size, n_max, 6)
    mask = torch.zeros(batch_size, n_max)
    order = torch.argsort(idx, stable=True)
    pos = torch.arange(len(idx)) - torch.cumsum(counts, 0)[idx[order]] + counts[idx[order]]
    keep = pos < n_max
    out[idx[order][keep], pos[keep]] = torch.cat([cls, box], 1)[order][keep]
    mask[idx[order][keep], pos[keep]] = 1.0
    return out, mask


def _cov(boxes):                                                     # metrics.py:178-195
    a = boxes[..., 2:3].pow(2) / 12
    b = boxes[..., 3:4].pow(2) / 12
    c = boxes[..., 4:5]
    cos, sin = c.cos(), c.sin()
    cos2, sin2 = cos.pow(2), sin.pow(2)
    return a * cos2 + b * sin2, a * sin2 + b * cos2, (a - b) * cos * sin


def probiou(obb1, obb2, eps: float = 1e-7):                          # metrics.py:198-233
    x1, y1 = obb1[..., 0:1], obb1[..., 1:2]
    x2, y2 = obb2[..., 0:1], obb2[..., 1:2]
    a1, b1, c1 = _cov(obb1)
    a2, b2, c2 = _cov(obb2)
    den = (a1 + a2) * (b1 + b2) - (c1 + c2).pow(2) + eps
    t1 = (((a1 + a2) * (y1 - y2).pow(2) + (b1 + b2) * (x1 - x2).pow(2)) / den) * 0.25
    t2 = (((c1 + c2) * (x2 - x1) * (y1 - y2)) / den) * 0.5
    t3 = (((a1 + a2) * (b1 + b2) - (c1 + c2).pow(2))
          / (4 * ((a1 * b1 - c1.pow(2)).clamp(0) * (a2 * b2 - c2.pow(2)).clamp(0)).sqrt() + eps) + eps).log() * 0.5
    bd = (t1 + t2 + t3).clamp(eps, 100.0)
    hd = (1.0 - (-bd).exp() + eps).sqrt()
    return 1 - hd


class OBBLossStatic:
    def __init__(self, model=None, tal_topk: int = 10, *, stride=None, nc=None, reg_max=None, hyp=None, device=None):
        """`OBBLossStatic(model)` reads the head geometry and the loss gains from a reference OBBModel as v8OBBLoss.__init__ does
        (loss.py:401-419, :859-868); the keyword form takes them explicitly (hyp: object with .box / .cls / .dfl)."""
        if model is not None:
            m = model.model[-1]                                      # the OBB head
            stride, nc, reg_max, hyp = m.stride, m.nc, m.reg_max, model.args
            device = next(model.parameters()).device
        self.hyp = hyp
        self.stride = [float(v) for v in (stride.tolist() if hasattr(stride, "tolist") else stride)]   # host floats: no device scalars below
        self.nc, self.reg_max = nc, reg_max
        self.no = nc + reg_max * 4
        self.device = torch.device(device)
        self.topk, self.alpha, self.beta, self.eps = tal_topk, 0.5, 6.0, 1e-9      # loss.py:862
        self.lambda_angular, self.lambda_reg = 0.5, 0.05                           # loss.py:866-868
        self.proj = torch.arange(reg_max, dtype=torch.float32, device=self.device)
        self._anchors = {}
        self._ws = None

    def _make_anchors(self, shapes):                                 # tal.py:333-345 (cached per feature-map geometry)
        key = tuple(shapes)
        if key not in self._anchors:
            pts, strides = [], []
            for (h, w), s in zip(shapes, self.stride):
                sx = torch.arange(w, device=self.device, dtype=torch.float32) + 0.5
                sy = torch.arange(h, device=self.device, dtype=torch.float32) + 0.5
                sy, sx = torch.meshgrid(sy, sx, indexing="ij")
                pts.append(torch.stack((sx, sy), -1).view(-1, 2))
                strides.append(torch.full((h * w, 1), s, dtype=torch.float32, device=self.device))
            a, st = torch.cat(pts), torch.cat(strides)
            imgsz = (shapes[0][0] * self.stride[0], shapes[0][1] * self.stride[0])                   # (h, w), loss.py:956
            scale = torch.tensor([imgsz[1], imgsz[0], imgsz[1], imgsz[0]], dtype=torch.float32, device=self.device)
            self._anchors[key] = (a, st, (a * st).contiguous(), imgsz, scale)
        return self._anchors[key]

    def _assign(self, scores, boxes, anc_px, labels, gts, mask):
        B, A, nc = scores.shape
        n = gts.shape[1]
        lib = _lib.load()
        need = lib.quan_rotated_tal_workspace_bytes(B, A, n)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=scores.device)
        t_boxes = torch.empty(B, A, 5, dtype=torch.float32, device=scores.device)
        t_scores = torch.empty(B, A, nc, dtype=torch.float32, device=scores.device)
        fg = torch.empty(B, A, dtype=torch.bool, device=scores.device)
        tgi = torch.empty(B, A, dtype=torch.int64, device=scores.device)
        check(lib.quan_rotated_tal_assign(scores.data_ptr(), boxes.data_ptr(), anc_px.data_ptr(), labels.data_ptr(), gts.data_ptr(),
                                          mask.data_ptr(), B, A, n, nc, self.topk, self.alpha, self.beta, self.eps, t_boxes.data_ptr(),
                                          t_scores.data_ptr(), fg.data_ptr(), tgi.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                          torch.cuda.current_stream(scores.device).cuda_stream), "quan_rotated_tal_assign")
        return t_boxes, t_scores, fg, tgi

    def __call__(self, preds, batch):
        feats, pred_angle = preds if isinstance(preds[0], list) else preds[1]
        B = pred_angle.shape[0]
        x = torch.cat([xi.view(B, self.no, -1) for xi in feats], 2).float()
        pred_distri, pred_scores = x.split((self.reg_max * 4, self.nc), 1)
        pred_scores = pred_scores.permute(0, 2, 1).contiguous()                    # [B, A, nc]
        pred_distri = pred_distri.permute(0, 2, 1).contiguous()                    # [B, A, 64]
        pred_angle = pred_angle.float().permute(0, 2, 1).contiguous()              # [B, A, 1]
        shapes = [tuple(f.shape[2:]) for f in feats]
        anchor_points, stride_tensor, anc_px, imgsz, scale = self._make_anchors(shapes)

        # targets (loss.py:959-968): scale xywh to pixels, drop boxes thinner than 2 px, mask padding
        tg, tmask = batch["targets"].float(), batch["target_mask"].float()
        gt_labels = tg[..., 0].contiguous()
        gt_xywh = tg[..., 1:5] * scale
        gt_bboxes = torch.cat([gt_xywh, tg[..., 5:6]], -1)
        # loss.py:963-964 filters on rw = w_n * imgsz[0], rh = h_n * imgsz[1] (height/width swapped exactly as in the reference)
        keep = ((tg[..., 3] * imgsz[0]) >= 2) & ((tg[..., 4] * imgsz[1]) >= 2)
        mask_gt = (tmask * keep * (gt_bboxes.sum(2) > 0)).contiguous()
        gt_bboxes = (gt_bboxes * mask_gt.unsqueeze(-1)).contiguous()

        # predicted boxes (loss.py:978, :1035-1050)
        b, a, c = pred_distri.shape
        dist = pred_distri.view(b, a, 4, c // 4).softmax(3).matmul(self.proj)
        lt, rb = dist.split(2, -1)
        cos, sin = torch.cos(pred_angle), torch.sin(pred_angle)
        xf, yf = ((rb - lt) / 2).split(1, -1)
        xy = torch.cat([xf * cos - yf * sin, xf * sin + yf * cos], -1) + anchor_points
        pred_bboxes = torch.cat([xy, lt + rb, pred_angle], -1)                     # [B, A, 5] in grid units

        with torch.no_grad():
            for_assigner = torch.cat([pred_bboxes[..., :4] * stride_tensor, pred_bboxes[..., 4:]], -1).contiguous()
            target_bboxes, target_scores, fg_mask, _ = self._assign(pred_scores.sigmoid().contiguous(), for_assigner, anc_px,
                                                                    gt_labels, gt_bboxes, mask_gt)
            target_bboxes = torch.cat([target_bboxes[..., :4] / stride_tensor, target_bboxes[..., 4:]], -1)
            fg = fg_mask.float()
            tss = target_scores.sum().clamp_min(1.0)
            weight = target_scores.sum(-1) * fg                                    # [B, A]
            n_fg = fg.sum().clamp_min(1.0)

        loss_cls = F.binary_cross_entropy_with_logits(pred_scores, target_scores, reduction="none").sum() / tss
        # box + DFL (loss.py:364-378); background anchors carry weight 0 and finite values (their target is box 0 of the image)
        # (background anchors are compared with themselves: finite value and gradient, weight 0 — the reference never evaluates them)
        tb = torch.where(fg_mask.unsqueeze(-1), target_bboxes, pred_bboxes.detach())
        iou = probiou(pred_bboxes, tb).squeeze(-1)
        loss_iou = ((1.0 - iou) * weight).sum() / tss
        x1y1 = target_bboxes[..., :2] - target_bboxes[..., 2:4] / 2
        x2y2 = target_bboxes[..., :2] + target_bboxes[..., 2:4] / 2
        ltrb = torch.cat((anchor_points - x1y1, x2y2 - anchor_points), -1).clamp(0, self.reg_max - 1 - 0.01)
        tl = ltrb.long()
        wl = (tl + 1) - ltrb
        logp = pred_distri.view(b, a, 4, c // 4).log_softmax(3)
        ce_l = -logp.gather(3, tl.unsqueeze(-1)).squeeze(-1)
        ce_r = -logp.gather(3, (tl + 1).unsqueeze(-1)).squeeze(-1)
        loss_dfl = ((ce_l * wl + ce_r * (1 - wl)).mean(-1) * weight).sum() / tss
        # quaternion angular loss (loss.py:870-921, :1008-1025): rotation about z -> q = (cos(t/2), 0, 0, sin(t/2))
        hp, ht = pred_bboxes[..., 4] / 2, target_bboxes[..., 4] / 2
        dot = (torch.cos(hp) * torch.cos(ht) + torch.sin(hp) * torch.sin(ht)).clamp(-1.0 + 1e-7, 1.0 - 1e-7)
        ang = 2.0 * torch.arccos(dot.abs())
        nsq = torch.cos(hp).pow(2) + torch.sin(hp).pow(2)
        loss_ang = (ang * weight).sum() / tss + self.lambda_reg * (((nsq - 1.0) ** 2) * fg).sum() / n_fg

        loss = torch.stack([loss_iou * self.hyp.box, loss_cls * self.hyp.cls, loss_dfl * self.hyp.dfl, loss_ang * self.lambda_angular])
        return loss.sum() * B, loss.detach()


# ---- fully fused criterion: decode -> assigner -> loss + gradients, 9 kernel launches ---------------------------------------------
class _ObbLossFusedFn(torch.autograd.Function):
    """(feat_0, feat_1, feat_2, pred_angle) -> (total, items).  The forward kernel already leaves d(total)/d(input) in tensors shaped
    and laid out like the inputs; backward scales them by the incoming gradient of `total`."""

    @staticmethod
    def forward(ctx, f0, f1, f2, angle, crit, targets, target_mask):
        feats = [_pixel_rows(f) for f in (f0, f1, f2)]
        angle = angle.contiguous()
        total, items, grads = crit._run(feats, angle, targets, target_mask)     # grads: dense [B,H,W,ld] x 3 + the angle gradient
        ctx.save_for_backward(*grads)
        ctx.widths = [f.shape[1] for f in feats]
        ctx.mark_non_differentiable(items)
        return total, items

    @staticmethod
    def backward(ctx, g_total, g_items):
        grads = ctx.saved_tensors
        out = torch._foreach_mul(list(grads), g_total.to(grads[0].dtype))
        views = [o[..., :n].permute(0, 3, 1, 2) for o, n in zip(out[:3], ctx.widths)]
        return views[0], views[1], views[2], out[3], None, None, None


def _pixel_rows(f: torch.Tensor) -> torch.Tensor:
    """A logical [B,N,H,W] head tensor whose memory is pixel rows of N contiguous values, `stride(3)` elements apart (channels-last,
    or the padded rows functional.qer_cat writes); anything else is made channels-last."""
    B, N, H, W = f.shape
    sb, sn, sh, sw = f.stride()
    if sn == 1 and sw >= N and sh == W * sw and sb == H * sh:
        return f
    return f.contiguous(memory_format=torch.channels_last)


class OBBLossFused(OBBLossStatic):
    """v8OBBLoss (ultralytics/utils/loss.py:853-1050) in nine launches of libquan_sm100.so: quan_obb_decode (sigmoid scores + decoded
    boxes for the assigner) -> quan_rotated_tal_assign -> quan_obb_loss_fwd_bwd (loss items AND the gradients w.r.t. the head outputs,
    read / written in the head's channels-last memory).  Same call interface and return values as OBBLossStatic / the reference."""

    def _run(self, feats, angle, targets, target_mask):
        import ctypes as C
        lib = _lib.load()
        dev = angle.device
        B = angle.shape[0]
        shapes = [tuple(f.shape[2:]) for f in feats]
        _, _, anc_px, imgsz, scale = self._make_anchors(shapes)
        A = anc_px.shape[0]
        dt = 1 if angle.dtype == torch.bfloat16 else 0
        if angle.dtype not in (torch.bfloat16, torch.float32) or any(f.dtype != angle.dtype for f in feats):
            raise RuntimeError("OBBLossFused: head outputs must all be float32 or all bfloat16")
        hw = (C.c_int32 * 6)(*[v for s in shapes for v in s])
        st = (C.c_float * 3)(*self.stride)
        fp = (C.c_void_p * 3)(*[f.data_ptr() for f in feats])
        lds = [f.stride(3) for f in feats]                       # row pitch per level (>= no: qer_cat pads rows to 16 bytes)
        ldp = (C.c_int32 * 3)(*lds)
        stream = torch.cuda.current_stream(dev).cuda_stream
        # targets (loss.py:959-968)
        tg, tmask = targets.float(), target_mask.float()
        gt_labels = tg[..., 0].contiguous()
        gt_bboxes = torch.cat([tg[..., 1:5] * scale, tg[..., 5:6]], -1)
        keep = ((tg[..., 3] * imgsz[0]) >= 2) & ((tg[..., 4] * imgsz[1]) >= 2)
        mask_gt = (tmask * keep * (gt_bboxes.sum(2) > 0)).contiguous()
        gt_bboxes = (gt_bboxes * mask_gt.unsqueeze(-1)).contiguous()
        pd_scores = torch.empty(B, A, self.nc, dtype=torch.float32, device=dev)
        pd_bboxes = torch.empty(B, A, 5, dtype=torch.float32, device=dev)
        check(lib.quan_obb_decode(fp, angle.data_ptr(), hw, st, B, self.nc, self.reg_max, ldp, pd_scores.data_ptr(), pd_bboxes.data_ptr(), dt,
                                  stream), "quan_obb_decode")
        t_boxes, t_scores, fg, _ = self._assign(pd_scores, pd_bboxes, anc_px, gt_labels, gt_bboxes, mask_gt)
        grads = [torch.empty((B, f.shape[2], f.shape[3], ld), dtype=f.dtype, device=dev) for f, ld in zip(feats, lds)] + [torch.empty_like(angle)]
        gp = (C.c_void_p * 3)(*[g.data_ptr() for g in grads[:3]])
        scratch = torch.empty(5, dtype=torch.float64, device=dev)
        items = torch.empty(4, dtype=torch.float32, device=dev)
        total = torch.empty((), dtype=torch.float32, device=dev)
        check(lib.quan_obb_loss_fwd_bwd(fp, angle.data_ptr(), hw, st, B, self.nc, self.reg_max, t_boxes.data_ptr(), t_scores.data_ptr(),
                                        fg.data_ptr(), float(self.hyp.box), float(self.hyp.cls), float(self.hyp.dfl), float(self.lambda_angular),
                                        ldp, gp, grads[3].data_ptr(), scratch.data_ptr(), items.data_ptr(), total.data_ptr(), dt, stream),
              "quan_obb_loss_fwd_bwd")
        return total, items, grads

    def __call__(self, preds, batch):
        feats, pred_angle = preds if isinstance(preds[0], list) else preds[1]
        return _ObbLossFusedFn.apply(feats[0], feats[1], feats[2], pred_angle, self, batch["targets"], batch["target_mask"])
