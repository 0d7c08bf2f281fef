"""The loss half of the training step on the GPU: quan_rotated_tal_assign and loss.OBBLossStatic against golden
vectors recorded from the real reference (tests/golden/make_tal_golden.py: utils/tal.py RotatedTaskAlignedAssigner, utils/loss.py
v8OBBLoss on CPU), and against the reference criterion run live on the same device at a realistic size."""
import types
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = np.load(Path(__file__).resolve().parent / "golden" / "tal.npz")


def _crit(nc, reg_max, hyp):
    from quan_ultralytics_b200.loss import OBBLossStatic
    return OBBLossStatic(stride=[8.0, 16.0, 32.0], nc=nc, reg_max=reg_max, hyp=types.SimpleNamespace(box=hyp[0], cls=hyp[1], dfl=hyp[2]),
                         device="cuda")


def test_assigner_matches_reference_golden():
    t = lambda k: torch.from_numpy(G["asg_" + k]).cuda()
    crit = _crit(5, 16, (7.5, 0.5, 1.5))
    tb, ts, fg, tgi = crit._assign(t("pd_scores").contiguous(), t("pd_bboxes").contiguous(), t("anc").contiguous(),
                                   t("gt_labels").squeeze(-1).contiguous(), t("gt_bboxes").contiguous(), t("mask_gt").squeeze(-1).contiguous())
    torch.cuda.synchronize()
    assert torch.equal(fg.cpu(), torch.from_numpy(G["asg_fg_mask"]))
    assert torch.equal(tgi.cpu(), torch.from_numpy(G["asg_target_gt_idx"]))
    torch.testing.assert_close(ts.cpu(), torch.from_numpy(G["asg_target_scores"]).float(), rtol=2e-4, atol=1e-7)
    torch.testing.assert_close(tb.cpu(), torch.from_numpy(G["asg_target_bboxes"]), rtol=0, atol=0)


def test_static_loss_matches_reference_golden():
    from quan_ultralytics_b200.loss import pad_targets
    Bz, S, nc, reg_max = (int(v) for v in G["loss_meta"])
    crit = _crit(nc, reg_max, G["loss_hyp"])
    ins = [torch.from_numpy(G[f"loss_in{i}"]).cuda().requires_grad_(True) for i in range(4)]
    batch = {k: torch.from_numpy(G["loss_" + k]) for k in ("batch_idx", "cls", "bboxes")}
    tg, tm = pad_targets(batch, Bz)
    total, items = crit((ins[:3], ins[3]), {"targets": tg.cuda(), "target_mask": tm.cuda()})
    grads = torch.autograd.grad(total, ins)
    torch.cuda.synchronize()
    np.testing.assert_allclose(items.cpu().numpy(), G["loss_items"], rtol=2e-4)
    np.testing.assert_allclose(float(total), float(G["loss_total"]), rtol=2e-4)
    for i, g in enumerate(grads):
        ref = G[f"loss_grad{i}"]
        err = np.abs(g.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err <= 2e-4, (i, err)


def test_static_loss_vs_live_reference_at_model_size():
    """16 x 1024^2-shaped head outputs (A = 21504 anchors, 40 boxes per image): the reference criterion on the same device."""
    from quan_ultralytics_b200 import refenv
    if refenv.find_reference() is None:
        pytest.skip("no reference tree (baseline/_ref)")
    import math
    from quan_ultralytics_b200 import workloads
    from quan_ultralytics_b200.loss import OBBLossStatic, pad_targets
    refenv.activate()
    from ultralytics.utils.loss import v8OBBLoss
    torch.manual_seed(0)
    model = workloads.build_yolo_obb("n", 15, "cuda", swapped=False)
    ref, ours = v8OBBLoss(model), OBBLossStatic(model)
    Bz, S = 4, 1024
    batch = workloads.synthetic_obb_batch(Bz, S, "cuda", seed=4)
    g = torch.Generator(device="cuda").manual_seed(1)
    feats = [(torch.randn(Bz, ref.no, S // s, S // s, device="cuda", generator=g) * 0.5).requires_grad_(True) for s in (8, 16, 32)]
    angle = ((torch.rand(Bz, 1, sum((S // s) ** 2 for s in (8, 16, 32)), device="cuda", generator=g) - 0.25) * math.pi).requires_grad_(True)
    t_ref, i_ref = ref(([f for f in feats], angle), {k: v.clone() for k, v in batch.items()})
    g_ref = torch.autograd.grad(t_ref, feats + [angle])
    tg, tm = pad_targets(batch, Bz)
    t_our, i_our = ours(([f for f in feats], angle), {"targets": tg.cuda(), "target_mask": tm.cuda()})
    g_our = torch.autograd.grad(t_our, feats + [angle])
    torch.cuda.synchronize()
    print(f"\nloss {float(t_our):.4f} vs {float(t_ref):.4f}; items {i_our.tolist()} vs {i_ref.tolist()}")
    torch.testing.assert_close(i_our, i_ref, rtol=5e-4, atol=1e-6)
    for a, b in zip(g_our, g_ref):
        err = float((a - b).abs().max() / b.abs().max())
        assert err <= 1e-3, err


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 3e-4), (torch.bfloat16, 2e-2)])
def test_fused_loss_matches_reference_golden(dtype, tol):
    """quan_obb_decode + quan_rotated_tal_assign + quan_obb_loss_fwd_bwd against the reference criterion's recorded loss and gradients."""
    from quan_ultralytics_b200.loss import OBBLossFused, pad_targets
    import types
    Bz, S, nc, reg_max = (int(v) for v in G["loss_meta"])
    h = G["loss_hyp"]
    crit = OBBLossFused(stride=[8.0, 16.0, 32.0], nc=nc, reg_max=reg_max, hyp=types.SimpleNamespace(box=h[0], cls=h[1], dfl=h[2]), device="cuda")
    ins = [torch.from_numpy(G[f"loss_in{i}"]).cuda().to(dtype) for i in range(4)]
    ins = [t.contiguous(memory_format=torch.channels_last).requires_grad_(True) if t.dim() == 4 else t.requires_grad_(True) for t in ins]
    batch = {k: torch.from_numpy(G["loss_" + k]) for k in ("batch_idx", "cls", "bboxes")}
    tg, tm = pad_targets(batch, Bz)
    total, items = crit((ins[:3], ins[3]), {"targets": tg.cuda(), "target_mask": tm.cuda()})
    grads = torch.autograd.grad(total, ins)
    torch.cuda.synchronize()
    np.testing.assert_allclose(items.cpu().numpy(), G["loss_items"], rtol=tol)
    np.testing.assert_allclose(float(total), float(G["loss_total"]), rtol=tol)
    for i, g in enumerate(grads):
        ref = G[f"loss_grad{i}"]
        err = np.abs(g.float().cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err <= (tol if dtype == torch.float32 else 5e-2), (i, err)


def test_fused_loss_vs_live_reference_at_model_size():
    from quan_ultralytics_b200 import refenv
    if refenv.find_reference() is None:
        pytest.skip("no reference tree (baseline/_ref)")
    import math
    from quan_ultralytics_b200 import workloads
    from quan_ultralytics_b200.loss import OBBLossFused, pad_targets
    refenv.activate()
    from ultralytics.utils.loss import v8OBBLoss
    torch.manual_seed(0)
    model = workloads.build_yolo_obb("n", 15, "cuda", swapped=False)
    ref, ours = v8OBBLoss(model), OBBLossFused(model)
    Bz, S = 4, 1024
    batch = workloads.synthetic_obb_batch(Bz, S, "cuda", seed=4)
    g = torch.Generator(device="cuda").manual_seed(1)
    feats = [(torch.randn(Bz, ref.no, S // s, S // s, device="cuda", generator=g) * 0.5).contiguous(memory_format=torch.channels_last)
             .requires_grad_(True) for s in (8, 16, 32)]
    angle = ((torch.rand(Bz, 1, sum((S // s) ** 2 for s in (8, 16, 32)), device="cuda", generator=g) - 0.25) * math.pi).requires_grad_(True)
    t_ref, i_ref = ref(([f for f in feats], angle), {k: v.clone() for k, v in batch.items()})
    g_ref = torch.autograd.grad(t_ref, feats + [angle])
    tg, tm = pad_targets(batch, Bz)
    t_our, i_our = ours(([f for f in feats], angle), {"targets": tg.cuda(), "target_mask": tm.cuda()})
    g_our = torch.autograd.grad(t_our, feats + [angle])
    torch.cuda.synchronize()
    print(f"\nfused loss {float(t_our):.4f} vs {float(t_ref):.4f}; items {i_our.tolist()} vs {i_ref.tolist()}")
    torch.testing.assert_close(i_our, i_ref, rtol=5e-4, atol=1e-6)
    errs = [float((a - b).abs().max() / b.abs().max()) for a, b in zip(g_our, g_ref)]
    print("max gradient errors relative to the largest entry (3 levels, angle):", [f"{e:.1e}" for e in errs])
    # fp32 on both sides; the ProbIoU gradient divides by sqrt(1 - exp(-bd)) ~ 0 for near-perfect boxes, where the two evaluation
    # orders (autograd's reverse mode vs the kernel's forward-mode duals) differ in the last bits of a large quotient
    assert max(errs) <= 3e-3, errs


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_fused_loss_staged_rows_match_direct_rows(dtype):
    """The loss / decode kernels have two forms: rows staged through shared memory (pitch of whole 32-bit words: the padded rows of the
    fused head tensor, 16-byte copies when the pitch allows) and direct per-thread row access (any pitch).  Same arithmetic per anchor:
    padded bf16 rows (pitch 80: staged, vector copies), unpadded bf16 rows (pitch 79: direct) and fp32 rows (staged, word copies) must agree."""
    import types
    from quan_ultralytics_b200 import workloads
    from quan_ultralytics_b200.loss import OBBLossFused, pad_targets
    nc, reg_max, Bz, S = 15, 16, 3, 256
    no = 4 * reg_max + nc
    crit = OBBLossFused(stride=[8.0, 16.0, 32.0], nc=nc, reg_max=reg_max, hyp=types.SimpleNamespace(box=7.5, cls=0.5, dfl=1.5), device="cuda")
    batch = workloads.synthetic_obb_batch(Bz, S, "cuda", seed=7)
    tg, tm = pad_targets(batch, Bz)
    tgt = {"targets": tg.cuda(), "target_mask": tm.cuda()}
    g = torch.Generator(device="cuda").manual_seed(5)
    base = [(torch.randn(Bz, no, S // s, S // s, device="cuda", generator=g) * 0.5).to(dtype) for s in (8, 16, 32)]
    angle = ((torch.rand(Bz, 1, sum((S // s) ** 2 for s in (8, 16, 32)), device="cuda", generator=g) - 0.25) * 3.14159).to(dtype)

    def run(pad):
        feats = []
        for f in base:
            if pad:
                big = torch.zeros(Bz, no + pad, f.shape[2], f.shape[3], device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)
                big[:, :no] = f
                v = big[:, :no]
                assert v.stride(3) == no + pad
            else:
                v = f.contiguous(memory_format=torch.channels_last)
            feats.append(v.detach().requires_grad_(True))
        a = angle.clone().requires_grad_(True)
        total, items = crit((feats, a), tgt)
        grads = torch.autograd.grad(total, feats + [a])
        torch.cuda.synchronize()
        return total, items, grads

    t0, i0, g0 = run(0)
    t1, i1, g1 = run(1)
    torch.testing.assert_close(i1, i0, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(t1, t0, rtol=1e-6, atol=1e-7)
    for a, b in zip(g1, g0):
        assert torch.equal(a.float(), b.float())


def test_assigner_vs_live_reference_giant_and_tiny_boxes():
    """Top-k edge cases against the reference's own RotatedTaskAlignedAssigner (tal.py:298-330) on the device: a box that covers more
    anchors than the kernel's compact candidate list holds (full-row scans), a box so small that fewer than k anchors have a positive
    metric (the remaining picks are the lowest-index zeros), a masked-out box, and ordinary boxes."""
    from quan_ultralytics_b200 import refenv
    if refenv.find_reference() is None:
        pytest.skip("no reference tree (baseline/_ref)")
    refenv.activate()
    from ultralytics.utils.tal import RotatedTaskAlignedAssigner
    from ultralytics.utils.tal import make_anchors
    torch.manual_seed(3)
    Bz, S, nc = 2, 1024, 15
    feats = [torch.zeros(Bz, 1, S // s, S // s, device="cuda") for s in (8, 16, 32)]
    anc, strides = make_anchors(feats, [8, 16, 32], 0.5)
    anc_px = (anc * strides).contiguous()
    A = anc_px.shape[0]
    pd_scores = torch.rand(Bz, A, nc, device="cuda")
    # predicted boxes near their anchors, random sizes / angles
    pd_bboxes = torch.cat([anc_px.unsqueeze(0).expand(Bz, -1, -1) + torch.randn(Bz, A, 2, device="cuda") * 4,
                           torch.rand(Bz, A, 2, device="cuda") * 200 + 8, (torch.rand(Bz, A, 1, device="cuda") - 0.25) * 3.14159], -1).contiguous()
    gt = torch.tensor([[[512., 512., 900., 880., 0.3], [100., 120., 5., 4., 0.1], [700., 300., 120., 60., 1.0], [0., 0., 0., 0., 0.]],
                       [[300., 640., 640., 700., -0.4], [900., 900., 30., 90., 0.7], [40., 40., 3., 3., 0.0], [512., 512., 64., 64., 0.2]]],
                      device="cuda")
    labels = torch.tensor([[[1.], [3.], [7.], [0.]], [[2.], [2.], [14.], [5.]]], device="cuda")
    mask = torch.tensor([[[1.], [1.], [1.], [0.]], [[1.], [1.], [1.], [1.]]], device="cuda")
    ref = RotatedTaskAlignedAssigner(topk=10, num_classes=nc, alpha=0.5, beta=6.0)
    _, tb_r, ts_r, fg_r, tgi_r = ref(pd_scores, pd_bboxes, anc_px, labels, gt, mask)
    crit = _crit(nc, 16, (7.5, 0.5, 1.5))
    tb, ts, fg, tgi = crit._assign(pd_scores, pd_bboxes, anc_px, labels.squeeze(-1).contiguous(), gt.contiguous(), mask.squeeze(-1).contiguous())
    torch.cuda.synchronize()
    assert torch.equal(fg, fg_r.bool())
    assert torch.equal(tgi[fg], tgi_r[fg_r.bool()])
    torch.testing.assert_close(ts, ts_r.float(), rtol=2e-4, atol=1e-7)
    torch.testing.assert_close(tb[fg], tb_r[fg_r.bool()], rtol=0, atol=0)
    assert int(fg.sum()) > 20


def test_assigner_full_scan_path_in_subprocess():
    """The assigner's top-k has two forms (compact candidate list / full-row scans; the second only when a box has more positive anchors
    than the list holds).  QUAN_TAL_FULLSCAN=1 forces the full scans — the switch is read once per process, so the assigner tests run
    again in a child process with it set."""
    import os
    import subprocess
    import sys
    if os.environ.get("QUAN_TAL_FULLSCAN") == "1":
        pytest.skip("already the child process")
    env = dict(os.environ, QUAN_TAL_FULLSCAN="1")
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-m", "gpu", "-k",
                        "assigner_matches_reference_golden or giant_and_tiny or fused_loss_matches_reference_golden"],
                       env=env, capture_output=True, text=True, timeout=600, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
