"""Golden vectors for the rotated task-aligned assigner and the OBB loss, from the REAL reference on CPU (fp32):
    python tests/golden/make_tal_golden.py  ->  tests/golden/tal.npz
 * assigner: ultralytics/utils/tal.py:298 RotatedTaskAlignedAssigner(topk=10, nc, alpha=0.5, beta=6.0) on random predictions whose
   boxes sit near their anchors (as decoded predictions do), with padded / masked ground truths;
 * loss: ultralytics/utils/loss.py:853 v8OBBLoss on random head outputs of a QUAN-YOLO11n-OBB model object (strides 8/16/32):
   total, items, and the gradients w.r.t. the three feature maps and the angle logits."""
import math
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from quan_ultralytics_b200 import refenv, workloads  # noqa: E402

refenv.activate(Path("/root/reference"))
from ultralytics.utils.loss import v8OBBLoss  # noqa: E402
from ultralytics.utils.tal import RotatedTaskAlignedAssigner  # noqa: E402

out = {}
g = torch.Generator().manual_seed(0)
# ---- assigner ---------------------------------------------------------------------------------------------------------------------
B, n, nc = 3, 7, 5
hw = [(16, 16), (8, 8), (4, 4)]
strides = [8, 16, 32]
anc = torch.cat([torch.stack(torch.meshgrid(torch.arange(h) + 0.5, torch.arange(w) + 0.5, indexing="ij")[::-1], -1).view(-1, 2) * s
                 for (h, w), s in zip(hw, strides)])
A = anc.shape[0]
pd_scores = torch.rand(B, A, nc, generator=g)
wh = torch.rand(B, A, 2, generator=g) * 40 + 6
pd_bboxes = torch.cat([anc[None] + (torch.rand(B, A, 2, generator=g) - 0.5) * 10, wh, torch.rand(B, A, 1, generator=g) * math.pi - math.pi / 4], -1)
gt_bboxes = torch.cat([torch.rand(B, n, 2, generator=g) * 100 + 14, torch.rand(B, n, 2, generator=g) * 50 + 8,
                       torch.rand(B, n, 1, generator=g) * math.pi - math.pi / 4], -1)
gt_labels = torch.randint(0, nc, (B, n, 1), generator=g).float()
mask_gt = torch.ones(B, n, 1)
mask_gt[0, 5:] = 0
mask_gt[2, :] = 0                                   # an image without boxes
gt_bboxes = gt_bboxes * mask_gt
asg = RotatedTaskAlignedAssigner(topk=10, num_classes=nc, alpha=0.5, beta=6.0)
tl, tb, ts, fg, tgi = asg(pd_scores, pd_bboxes, anc, gt_labels, gt_bboxes, mask_gt)
for k, v in dict(pd_scores=pd_scores, pd_bboxes=pd_bboxes, anc=anc, gt_labels=gt_labels, gt_bboxes=gt_bboxes, mask_gt=mask_gt,
                 target_bboxes=tb, target_scores=ts, fg_mask=fg, target_gt_idx=tgi).items():
    out["asg_" + k] = v.numpy()

# ---- loss ---------------------------------------------------------------------------------------------------------------------------
torch.manual_seed(0)
model = workloads.build_yolo_obb("n", 15, "cpu", swapped=False)
crit = v8OBBLoss(model)
Bz, S = 2, 128
batch = workloads.synthetic_obb_batch(Bz, S, "cpu", boxes_per_image=6, seed=2)
batch["bboxes"][3, 2:4] = 0.001                      # a box thinner than 2 px: filtered by loss.py:963-964
feats = [torch.randn(Bz, crit.no, S // s, S // s, generator=g).requires_grad_(True) for s in (8, 16, 32)]
angle = ((torch.rand(Bz, 1, sum((S // s) ** 2 for s in (8, 16, 32)), generator=g) - 0.25) * math.pi).requires_grad_(True)
total, items = crit((feats, angle), batch)
grads = torch.autograd.grad(total, feats + [angle])
out["loss_total"], out["loss_items"] = total.detach().numpy(), items.numpy()
for i, (f, gr) in enumerate(zip(feats + [angle], grads)):
    out[f"loss_in{i}"], out[f"loss_grad{i}"] = f.detach().numpy(), gr.numpy()
for k in ("batch_idx", "cls", "bboxes"):
    out["loss_" + k] = batch[k].numpy()
out["loss_meta"] = np.array([Bz, S, crit.nc, crit.reg_max])
out["loss_hyp"] = np.array([crit.hyp.box, crit.hyp.cls, crit.hyp.dfl])
np.savez_compressed(ROOT / "tests" / "golden" / "tal.npz", **out)
print({k: v.shape for k, v in out.items()}, "fg anchors:", int(fg.sum()), "loss", float(total), items)
