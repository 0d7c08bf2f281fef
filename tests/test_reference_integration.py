"""INTEGRATION.md level 1/2 against the REAL reference checkout (only where /root/reference exists, i.e. the build
container; the GPU box has no reference): the class swap makes the reference's own model builders instantiate the
B200 modules with identical state-dict keys and shapes, and the extension shim is what `import quaternion_ops` finds."""
import os
import sys
import tempfile
import types
from pathlib import Path

import pytest
import torch

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference checkout not present on this box")


@pytest.fixture(scope="module")
def reference():
    os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp(prefix="yolo_cfg_"))
    sys.dont_write_bytecode = True
    for name in ("matplotlib", "matplotlib.pyplot", "thop"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "matplotlib":
                m.use = lambda *a, **k: None
                m.rcParams = {}
                m.rc = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    for p in (str(REF), str(REF / "classification")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ultralytics.nn.modules.conv as uconv
    import quaternion.qconv as cconv
    return uconv, cconv


def _keys(m):
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def test_layer_state_dicts_interchange(reference):
    import quan_ultralytics_b200 as Q
    uconv, cconv = reference
    for ref_cls, ours, args, kw in [
        (uconv.QConv2D, Q.QConv2D, (64, 128, 3), dict(stride=2, padding=1, groups=2, bias=True)),
        (uconv.QConv2D, Q.QConv2D, (3, 16, 3), dict(stride=2, padding=1, bias=False)),
        (cconv.QConv2D, Q.QConv2D_B, (32, 32, 3), dict(padding=1, bias=True)),
        (uconv.IQBN, Q.IQBN, (64,), {}),
        (cconv.IQBN, Q.IQBN, (32,), {}),
        (uconv.Conv, Q.Conv, (64, 64, 3, 1), {}),
        (uconv.DWConv, Q.DWConv, (64, 64, 3), {}),
    ]:
        r, o = ref_cls(*args, **kw), ours(*args, **kw)
        assert _keys(r) == _keys(o), ref_cls
        o.load_state_dict(r.state_dict())          # checkpoints move across unchanged
        for k, v in r.state_dict().items():
            assert torch.equal(o.state_dict()[k], v)
        if hasattr(r, "padding"):
            assert tuple(r.padding) == tuple(o.padding) and tuple(r.stride) == tuple(o.stride)


def test_init_distribution_matches_reference(reference):
    import quan_ultralytics_b200 as Q
    uconv, _ = reference
    torch.manual_seed(5)
    r = uconv.QConv2D(128, 128, 3, bias=True)
    torch.manual_seed(5)
    o = Q.QConv2D(128, 128, 3, bias=True)
    for n in ("weight_r", "weight_i", "weight_j", "weight_k", "bias_r"):      # same RNG stream, same init rule
        assert torch.equal(getattr(r, n), getattr(o, n)), n


def test_class_swap_builds_reference_graphs_with_b200_modules(reference):
    import quan_ultralytics_b200 as Q
    import quan_ultralytics_b200.install as qi
    uconv, cconv = reference
    import ultralytics.nn.tasks as tasks
    import yaml
    import ultralytics.nn.modules.block as ublock
    import models.blocks.quaternion_blocks as cblocks
    import models.quaternion_models as cmodels
    import ultralytics.nn.modules as umods
    import ultralytics.nn.modules.head as uhead
    saved = {(m, n): getattr(m, n) for m in (uconv, tasks, ublock, cblocks, cmodels, umods, uhead)
             for n in ("QConv2D", "IQBN", "Conv", "DWConv", "QUpsample", "QuaternionMaxPool", "QER") if hasattr(m, n)}
    saved_c = (cconv.QConv2D, cconv.IQBN)
    try:
        cfg = yaml.safe_load((REF / "ultralytics/cfg/models/11/yolo11-obb-quan.yaml").read_text())
        cfg["scale"] = "n"
        cfg["nc"] = 15
        import copy
        ref_model, _ = tasks.parse_model(copy.deepcopy(cfg), ch=3, verbose=False)
        done = qi.install(ultralytics=True, classification=True)
        assert "QConv2D" in done["ultralytics.nn.modules.conv"] and "Conv" in done["ultralytics.nn.tasks"]
        our_model, _ = tasks.parse_model(copy.deepcopy(cfg), ch=3, verbose=False)
        assert _keys(ref_model) == _keys(our_model)                       # yaml loads unchanged, same parameters
        n_q = sum(isinstance(m, Q.QConv2D) for m in our_model.modules())
        n_bn = sum(isinstance(m, Q.IQBN) for m in our_model.modules())
        n_up = sum(isinstance(m, Q.QUpsample) for m in our_model.modules())
        # blocks that import QConv2D/IQBN by name at module scope keep building (87 convs / 84 norms in YOLO11n)
        assert n_q >= 60 and n_bn >= 60 and n_up == 2, (n_q, n_bn, n_up)
        # QSPPF (block.py:270-302) pools with the B200 QuaternionMaxPool; so does the Q-ResNet-34 stem
        assert sum(isinstance(m, Q.QuaternionMaxPool) for m in our_model.modules()) >= 1
        assert isinstance(cmodels.create_qrn34_imagenet(10).maxpool, Q.QuaternionMaxPool)
        assert sum(isinstance(m, Q.QER) for m in our_model.modules()) == 9        # 3 scales x (box, cls, angle) extractions
        our_model.load_state_dict(ref_model.state_dict())
        assert cconv.QConv2D is Q.QConv2D_B and cconv.IQBN is Q.IQBN
    finally:
        for (m, n), v in saved.items():
            setattr(m, n, v)
        cconv.QConv2D, cconv.IQBN = saved_c


def test_extension_shim_is_importable_as_quaternion_ops(reference):
    import quan_ultralytics_b200.install as qi
    old = sys.modules.pop("quaternion_ops", None)
    try:
        shim = qi.install_extension_shim("A")
        import quaternion_ops
        assert quaternion_ops is shim and quaternion_ops.get_mixing() == "A"
        for fn in ("qconv_forward", "qconv_backward", "iqbn_forward"):       # quaternion_ops_py.cpp:132-165
            assert callable(getattr(quaternion_ops, fn))
        with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
            quaternion_ops.iqbn_forward(torch.zeros(1, 1, 2, 2, 4), torch.ones(1, 4), torch.zeros(1, 4),
                                        torch.zeros(1, 4), torch.ones(1, 4), 1e-5)
    finally:
        quaternion_ops.set_mixing("B")
        sys.modules.pop("quaternion_ops", None)
        if old is not None:
            sys.modules["quaternion_ops"] = old


def test_qer_matches_the_reference_head_module(reference):
    """ultralytics/nn/modules/head.py:26-47 QER against ours: same state dict, same outputs and gradients for the
    reference's BCHWQ input and for the tensor-core layout (where ours skips the permute copy)."""
    import quan_ultralytics_b200 as Q
    import ultralytics.nn.modules.head as uhead
    torch.manual_seed(4)
    assert uhead.QER is not Q.QER and uhead.QER.__module__ == "ultralytics.nn.modules.head"
    r = uhead.QER(64, 15, 1).double()
    o = Q.QER(64, 15, 1).double()
    assert _keys(r) == _keys(o)
    o.load_state_dict(r.state_dict())
    x = torch.randn(2, 16, 6, 5, 4, dtype=torch.float64)
    xr = x.clone().requires_grad_(True)
    yr = r(xr)
    dy = torch.randn_like(yr)
    yr.backward(dy)
    for fmt in (torch.contiguous_format, torch.channels_last_3d):
        o.zero_grad(set_to_none=True)
        xo = x.clone().contiguous(memory_format=fmt).requires_grad_(True)
        yo = o(xo)
        yo.backward(dy)
        torch.testing.assert_close(yo, yr, rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(xo.grad, xr.grad, rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(o.output_proj.weight.grad, r.output_proj.weight.grad, rtol=1e-12, atol=1e-12)

