"""The reference's OWN CUDA extension (ultralytics/nn/cuda/quaternion_ops.cu, compiled by oracle/build_ref_ext.py into
oracle/_ref/quaternion_ops.so and shipped prebuilt to the GPU box) against our drop-in for the same module API
(quan_ultralytics_b200/quaternion_ops.py) on identical fp32 inputs: this pins parity on the real reference code on the
GPU, not only on the CPU golden vectors.  Skipped where the prebuilt extension is absent.  The extension computes M_B.
Known reference defect left out of the comparison: its bias gradient is sum(dY_r) instead of the autograd-correct
sum((M^T dY)_r), so cases run with bias_defined=False for the backward."""
import pytest
import torch

from oracle import build_ref_ext

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref():
    mod = build_ref_ext.load()
    if mod is None:
        pytest.skip("oracle/_ref/quaternion_ops.so not built (python oracle/build_ref_ext.py in the build container)")
    return mod


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


# (B, C_i, C_o, H, W, k, stride, pad, dil, groups, bias)
CASES = [(2, 8, 8, 12, 10, 3, 1, 1, 1, 1, False), (3, 4, 8, 13, 11, 3, 2, 1, 1, 1, True), (2, 16, 8, 9, 9, 1, 1, 0, 1, 1, False),
         (2, 8, 8, 10, 10, 3, 1, 2, 2, 2, True), (1, 64, 64, 16, 16, 3, 1, 1, 1, 1, False)]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(str(v) for v in c))
def test_qconv_forward_backward_match_the_reference_extension(ref, case):
    from quan_ultralytics_b200 import quaternion_ops as ours
    ours.set_mixing("B")
    B, ci, co, H, W, k, s, p, d, g, bias = case
    torch.manual_seed(17)
    x = torch.randn(B, ci, H, W, 4, device=DEV)
    w = [torch.randn(co, ci // g, k, k, device=DEV) / (ci * k * k) ** 0.5 for _ in range(4)]
    b = torch.randn(co, device=DEV) if bias else None
    args = ([s, s], [p, p], [d, d], g)
    y_ref = ref.qconv_forward(x, *w, b, None, None, None, *args)
    y = ours.qconv_forward(x, *w, b, None, None, None, *args)
    assert y.shape == y_ref.shape and rel(y, y_ref) <= 1e-5
    dy = torch.randn_like(y_ref)
    g_ref = ref.qconv_backward(dy, x, *w, False, *args)
    g_our = ours.qconv_backward(dy, x, *w, False, *args)
    for a, r in zip(g_our[:5], g_ref[:5]):                      # dX, dW_r, dW_i, dW_j, dW_k
        assert a.shape == r.shape and rel(a, r) <= 2e-5


def test_large_layers_take_the_tensor_core_path_within_the_tf32_budget(ref):
    """>= 2^18 elements and a tensor-core shape: the shim converts BCHWQ -> BHWQC, runs the tcgen05 engine (tf32 MMA on fp32
    tensors) and converts back — same contract, BASELINE.json's 1e-3 instead of exact fp32."""
    from quan_ultralytics_b200 import ops
    from quan_ultralytics_b200 import quaternion_ops as ours
    ours.set_mixing("B")
    torch.manual_seed(23)
    x = torch.randn(4, 64, 32, 32, 4, device=DEV)
    w = [torch.randn(64, 64, 3, 3, device=DEV) / 24.0 for _ in range(4)]
    args = ([1, 1], [1, 1], [1, 1], 1)
    assert ours._tensor_core_ok(x, w[0], *args, (0, 1, 2))
    y_ref = ref.qconv_forward(x, *w, None, None, None, None, *args)
    y = ours.qconv_forward(x, *w, None, None, None, None, *args)
    assert y.is_contiguous() and ops.layout_of(y) == ops.LAYOUT_BCHWQ and 1e-6 < rel(y, y_ref) <= 1e-3
    dy = torch.randn_like(y_ref)
    g_ref = ref.qconv_backward(dy, x, *w, False, *args)
    g_our = ours.qconv_backward(dy, x, *w, False, *args)
    assert g_our[0].is_contiguous()
    for a, r in zip(g_our[:5], g_ref[:5]):
        assert a.shape == r.shape and rel(a, r) <= 1e-3
    ours.set_fast_layout(False)
    try:
        assert rel(ours.qconv_forward(x, *w, None, None, None, None, *args), y_ref) <= 1e-5      # exact-fp32 path on request
    finally:
        ours.set_fast_layout(True)


def test_iqbn_forward_matches_the_reference_extension(ref):
    from quan_ultralytics_b200 import quaternion_ops as ours
    torch.manual_seed(3)
    C = 12
    x = torch.randn(3, C, 7, 9, 4, device=DEV)
    gamma, beta = torch.rand(C, 4, device=DEV) + 0.5, torch.randn(C, 4, device=DEV)
    mean, var = torch.randn(C, 4, device=DEV), torch.rand(C, 4, device=DEV) + 0.2
    y_ref = ref.iqbn_forward(x, gamma, beta, mean, var, 1e-5)
    y = ours.iqbn_forward(x, gamma, beta, mean, var, 1e-5)
    assert rel(y, y_ref) <= 1e-5
