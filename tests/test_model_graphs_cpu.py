"""The reference's REAL model graphs with the B200 classes installed, on CPU: the kernels behind `ops` are replaced by the torch
emulation in tests/emu_ops.py (test infrastructure, same signatures / memory formats), everything above the C ABI is the product:
class swap, state dicts, autograd wiring, the fused `Conv` node, BHWQC (channels_last_3d) tensors flowing through the reference's own
blocks (C3k2 chunk / cat, QSPPF, QC2PSA / QAttention reshape-permute, QER, OBB head views, v8OBBLoss views).  Loss and every
parameter gradient are compared with the untouched reference (its PyTorch path, batch-statistics IQBN).  The same comparison runs on
the real kernels in tests/test_gpu_models.py."""
import pytest
import torch

from quan_ultralytics_b200 import install as qi
from quan_ultralytics_b200 import refenv, workloads

pytestmark = pytest.mark.skipif(refenv.find_reference() is None, reason="no reference tree (baseline/_ref or /root/reference)")


def _grads(model):
    return {n: p.grad.detach().double().cpu().clone() for n, p in model.named_parameters() if p.grad is not None}


def _rel(a, b, floor=0.0):
    return float((a - b).abs().max() / b.abs().max().clamp_min(max(floor, 1e-30)))


def _worst(g_our, g_ref):
    """Largest per-tensor relative error; tensors whose reference gradient is mathematically zero (e.g. the beta of an IQBN that feeds
    another batch-norm: |g| ~ 1e-8 of rounding noise) are measured against 1e-3 of the largest gradient instead of their own noise."""
    floor = 1e-3 * max(float(g.abs().max()) for g in g_ref.values())
    return max((_rel(g_our[n], g_ref[n], floor), n) for n in g_ref)


def _yolo_step(model, batch):
    model.train()
    model.zero_grad(set_to_none=True)
    loss, items = model(batch)
    loss.backward()
    return float(loss), items.double().cpu(), _grads(model)


@pytest.fixture()
def emu():
    from tests import emu_ops
    with emu_ops.emulated() as o:
        yield o
    qi.uninstall()


@pytest.mark.parametrize("size", [128])
def test_yolo11n_obb_quan_graph_matches_reference(emu, size):
    torch.manual_seed(0)
    ref = workloads.build_yolo_obb("n", 15, "cpu", swapped=False)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    ours = workloads.build_yolo_obb("n", 15, "cpu", swapped=True)
    import quan_ultralytics_b200 as Q
    assert sum(isinstance(m, Q.QConv2D) for m in ours.modules()) == 87
    assert sum(isinstance(m, Q.IQBN) for m in ours.modules()) == 84
    ours.load_state_dict(sd)
    batch = workloads.synthetic_obb_batch(2, size, "cpu", boxes_per_image=6, seed=3)
    l_ref, it_ref, g_ref = _yolo_step(ref, {k: v.clone() for k, v in batch.items()})
    l_our, it_our, g_our = _yolo_step(ours, {k: v.clone() for k, v in batch.items()})
    assert abs(l_our - l_ref) <= 1e-4 * abs(l_ref), (l_our, l_ref)
    torch.testing.assert_close(it_our, it_ref, rtol=1e-4, atol=1e-6)
    assert set(g_our) == set(g_ref)                      # the same 3 parameters stay without gradient
    worst = _worst(g_our, g_ref)
    assert worst[0] <= 1e-3, worst
    # running statistics moved the same way (in place here, re-assigned in the reference)
    for (n, b_o), (_, b_r) in zip(ours.named_buffers(), ref.named_buffers()):
        if n.endswith("running_mean") or n.endswith("running_var"):
            torch.testing.assert_close(b_o, b_r, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name,B,size,nc", [("qwrn16_2", 8, 32, 10), ("qresnet34", 2, 64, 10)])
def test_classification_graphs_match_reference(emu, name, B, size, nc):
    torch.manual_seed(0)
    ref = workloads.build_classifier(name, nc, "cpu", swapped=False)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    ours = workloads.build_classifier(name, nc, "cpu", swapped=True)
    ours.load_state_dict(sd)
    x, y = workloads.synthetic_classification_batch(B, size, nc)
    out = []
    for m in (ref, ours):
        m.train()
        torch.manual_seed(7)              # the Q-ResNet blocks draw dropout masks (fresh contiguous tensors: layout-independent)
        logits = m(x.clone())
        loss = torch.nn.functional.cross_entropy(logits, y)
        loss.backward()
        out.append((float(loss), _grads(m)))
    (l_ref, g_ref), (l_our, g_our) = out
    assert abs(l_our - l_ref) <= 1e-4 * abs(l_ref), (l_our, l_ref)
    assert set(g_our) == set(g_ref)
    worst = _worst(g_our, g_ref)
    assert worst[0] <= 1e-3, worst


def test_pooled_batch_counters_count_like_the_reference(emu):
    """modules.pool_batch_counters (workloads.build_* call it): every IQBN keeps its `num_batches_tracked` buffer and state-dict
    entry (conv.py:516, :563) but all of them live in one tensor that a forward pre-hook of the model bumps once per training
    forward; eval forwards do not count; load_state_dict writes through the views."""
    import quan_ultralytics_b200 as Q
    torch.manual_seed(0)
    ref = workloads.build_classifier("qwrn16_2", 10, "cpu", swapped=False)
    ours = workloads.build_classifier("qwrn16_2", 10, "cpu", swapped=True)
    assert sorted(ours.state_dict()) == sorted(ref.state_dict())
    bns = [m for m in ours.modules() if isinstance(m, Q.IQBN)]
    assert len(bns) == 13 and all(m._nbt_pooled for m in bns)
    x, _ = workloads.synthetic_classification_batch(4, 32, 10)
    ours.train()
    ours(x)
    ours(x)
    ref.train()
    ref(x)
    ref(x)
    ours.eval()
    with torch.no_grad():
        ours(x)
    want = {k: int(v) for k, v in ref.state_dict().items() if k.endswith("num_batches_tracked")}
    got = {k: int(v) for k, v in ours.state_dict().items() if k.endswith("num_batches_tracked")}
    assert got == want and set(got.values()) == {2}
    sd = {k: (torch.tensor(7) if k.endswith("num_batches_tracked") else v) for k, v in ours.state_dict().items()}
    ours.load_state_dict(sd)
    assert int(ours._quan_nbt_pool.min()) == 7 and int(bns[3].num_batches_tracked) == 7
