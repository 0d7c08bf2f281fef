"""QuaternionMaxPool (ultralytics/nn/modules/block.py:85-109): oracle pinned to the reference's own
outputs (tests/golden/qpool.npz, made by tests/golden/make_pool_golden.py), CUDA path against the oracle.  Pure selection
and routing, so every comparison is exact (bit-exact values; gradients exact up to fp32/bf16 summation of the few
windows that share a maximum)."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import quan_oracle as O

GOLD = np.load(Path(__file__).parent / "golden" / "qpool.npz")
CASES = ["sppf_k5", "stem_k3s2", "default_k2"]
DEV = "cuda:0"


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference(case):
    k, s, p = (int(v) for v in GOLD[f"{case}/cfg"])
    x = GOLD[f"{case}/x"]
    y, arg = O.qmaxpool_fwd(x, k, s, p)
    np.testing.assert_array_equal(y, GOLD[f"{case}/y"])
    np.testing.assert_allclose(O.qmaxpool_bwd(GOLD[f"{case}/dy"], arg, x.shape[2:4]), GOLD[f"{case}/dx"], rtol=0, atol=1e-12)


def test_oracle_tie_rule_is_first_maximum():
    x = np.zeros((1, 1, 4, 4, 4))
    y, arg = O.qmaxpool_fwd(x, 2, 2, 0)
    assert (arg[0, 0, :, :, 0] == np.array([[0, 2], [8, 10]])).all() and (y == 0).all()
    dx = O.qmaxpool_bwd(np.ones_like(y), arg, (4, 4))
    assert dx[0, 0, 0, 0, 0] == 1 and dx[0, 0, 0, 1, 0] == 0 and dx.sum() == 16


def test_module_mirrors_reference_constructor():
    import quan_ultralytics_b200 as Q
    m = Q.QuaternionMaxPool(kernel_size=5, stride=1, padding=2)
    assert (m.kernel_size, m.stride, m.padding) == (5, 1, 2) and len(list(m.parameters())) == 0
    assert (Q.QuaternionMaxPool().kernel_size, Q.QuaternionMaxPool().stride, Q.QuaternionMaxPool().padding) == (2, 2, 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 4, 8, 8, 4))
    with pytest.raises(AssertionError):
        m(torch.randn(1, 16, 8, 8))


def test_c_abi_argument_errors():
    from quan_ultralytics_b200 import _lib
    lib = _lib.load()
    assert lib.quan_qmaxpool_fwd(None, None, None, 1, 1, 4, 4, 2, 2, 2, 2, 0, 0, 0, 0, None) == -1
    assert lib.quan_qmaxpool_fwd(1, 1, None, 1, 1, 4, 4, 2, 2, 2, 2, 2, 2, 0, 0, None) == -2      # pad > kernel / 2
    assert b"padding" in lib.quan_last_error()
    assert lib.quan_qmaxpool_fwd(1, 1, None, 1, 1, 2, 2, 5, 5, 1, 1, 0, 0, 0, 0, None) == -2      # window > input
    assert lib.quan_qmaxpool_bwd(1, None, 1, 1, 1, 4, 4, 2, 2, 2, 2, 0, 0, 0, 0, None) == -1


def _to_dev(a, dtype, layout):
    from quan_ultralytics_b200 import ops
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV, dtype)
    if layout == ops.LAYOUT_BHWQC:
        t = t.contiguous(memory_format=torch.channels_last_3d)
    return t


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_cuda_matches_golden(case, layout, dtype):
    import quan_ultralytics_b200 as Q
    k, s, p = (int(v) for v in GOLD[f"{case}/cfg"])
    x = _to_dev(GOLD[f"{case}/x"], dtype, layout).requires_grad_(True)      # halves: exact in bf16
    y = Q.QuaternionMaxPool(k, s, p)(x)
    assert torch.equal(y.double().cpu(), torch.from_numpy(GOLD[f"{case}/y"]))
    dy = _to_dev(GOLD[f"{case}/dy"], dtype, layout)
    y.backward(dy)
    _, arg = O.qmaxpool_fwd(GOLD[f"{case}/x"], k, s, p)
    ref = O.qmaxpool_bwd(dy.double().cpu().numpy(), arg, x.shape[2:4])
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    assert float(np.abs(x.grad.double().cpu().numpy() - ref).max()) <= tol * float(np.abs(ref).max())
    with torch.no_grad():                                                     # inference form: no index tensor
        assert torch.equal(Q.QuaternionMaxPool(k, s, p)(x.detach()), y.detach())


# (B, C, H, W, k, s, p): ragged sizes, every vector width (4C = 4 .. 8-aligned), the model shapes scaled down
RANDOM_CASES = [(2, 1, 9, 7, 3, 2, 1), (3, 5, 13, 11, 5, 1, 2), (2, 32, 16, 16, 5, 1, 2), (2, 16, 28, 28, 3, 2, 1),
                (1, 6, 10, 9, 2, 2, 0), (2, 3, 7, 8, (3, 2), (2, 1), (1, 0)),
                # >= 2048 rows of >= 128 vectors in BHWQC: the row-tiled kernels (several rows per block), images change mid-block
                (80, 16, 56, 40, 3, 2, 1), (70, 8, 31, 33, 5, 1, 2)]


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("layout", [0, 1])
@pytest.mark.parametrize("cfg", RANDOM_CASES, ids=lambda c: "x".join(str(v) for v in c[:4]) + f"_k{c[4]}")
def test_cuda_matches_oracle_random(cfg, layout, dtype):
    from quan_ultralytics_b200 import functional as QF
    B, C, H, W, k, s, p = cfg
    rng = np.random.default_rng(hash((B, C, H, W)) % 2**31)
    xn = rng.standard_normal((B, C, H, W, 4))
    xn[0, 0, :2, :2, :] = -np.inf                                            # an all -inf window keeps its first element
    x = _to_dev(xn, dtype, layout).requires_grad_(True)
    xr = x.detach().double().cpu().numpy()                                    # what the kernel sees (bf16-rounded: ties)
    y = QF.qmaxpool(x, k, s, p)
    y_ref, arg = O.qmaxpool_fwd(xr, k, s, p)
    assert torch.equal(y.double().cpu(), torch.from_numpy(y_ref))
    dy = _to_dev(rng.standard_normal(y_ref.shape), dtype, layout)
    y.backward(dy)
    ref = O.qmaxpool_bwd(dy.double().cpu().numpy(), arg, (H, W))
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    assert float(np.abs(x.grad.double().cpu().numpy() - ref).max()) <= tol * float(np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_nan_propagates_and_takes_the_gradient(dtype):
    """nn.MaxPool2d: a NaN in the window is the result, and the LAST NaN of the row-major scan receives the gradient."""
    from quan_ultralytics_b200 import functional as QF
    torch.manual_seed(1)
    x = torch.randn(1, 2, 6, 6, 4, device=DEV).to(dtype)
    x[0, 0, 1, 1, 0] = float("nan")
    x[0, 0, 2, 3, 0] = float("nan")
    x = x.contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    y = QF.qmaxpool(x, 3, 1, 1)
    y.backward(torch.ones_like(y))
    xr = x.detach().float().cpu().requires_grad_(True)
    yr = torch.stack([torch.nn.functional.max_pool2d(xr[..., q], 3, 1, 1) for q in range(4)], -1)
    yr.backward(torch.ones_like(yr))
    assert torch.equal(torch.isnan(y).cpu(), torch.isnan(yr)) and bool(torch.isnan(y).any())
    assert torch.equal(torch.nan_to_num(y.float().cpu()), torch.nan_to_num(yr.detach()))
    assert torch.equal(x.grad.float().cpu(), xr.grad)


@pytest.mark.gpu
def test_full_size_properties():
    """Model-size checks without the oracle: the Q-ResNet-34 stem pool (k3 s2 p1 on 112^2) and QSPPF's k5 s1 p2."""
    from quan_ultralytics_b200 import functional as QF
    torch.manual_seed(0)
    for shape, (k, s, p) in [((64, 16, 112, 112, 4), (3, 2, 1)), ((16, 32, 32, 32, 4), (5, 1, 2))]:
        x = torch.randn(shape, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
        y = QF.qmaxpool(x, k, s, p)
        # idempotence of the window maximum: pooling a constant-per-window upsample is the identity; monotone: y >= centre tap
        if s == 1:
            assert bool((y >= x.detach()).all())
        ref = torch.stack([torch.nn.functional.max_pool2d(x.detach()[..., q].float(), k, s, p) for q in range(4)], -1)
        assert torch.equal(y.float(), ref)
        y.backward(torch.ones_like(y))
        # every output element routes exactly one unit of gradient
        assert float(x.grad.float().sum()) == pytest.approx(y.numel(), rel=1e-3)
