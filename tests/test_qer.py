"""QER, the quaternion -> real extraction of the detection heads (ultralytics/nn/modules/head.py:26-47): the numpy
oracle pinned to the reference's own outputs (tests/golden/qer.npz, made by tests/golden/make_qer_golden.py from head.QER in fp64), the
CUDA kernels (csrc/qer.cu, through the C ABI) against the oracle and the golden vectors.  Tolerances: fp32 kernels are exact-fp32 FMA
(1e-5); bf16 kernels contract bf16 operands with fp32 accumulation and round the result to bf16 — compared with the oracle evaluated on
the same bf16-rounded operands, 1e-2 (BASELINE.json's bf16 bound) of the largest magnitude."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import quan_oracle as O

GOLD = np.load(Path(__file__).parent / "golden" / "qer.npz")
CASES = ["box_64_64", "cls_64_15", "angle_16_1", "s_cls_128_15"]
DEV = "cuda:0"


def _g(case, k):
    return GOLD[f"{case}/{k}"]


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference(case):
    x, w, b = _g(case, "x"), _g(case, "w"), _g(case, "b")
    np.testing.assert_allclose(O.qer_fwd(x, w, b), _g(case, "y"), rtol=0, atol=1e-12)
    dx, dw, db = O.qer_bwd(_g(case, "dy"), x, w)
    np.testing.assert_allclose(dx, _g(case, "dx"), rtol=0, atol=1e-12)
    np.testing.assert_allclose(dw, _g(case, "dw"), rtol=0, atol=1e-11)
    np.testing.assert_allclose(db, _g(case, "db"), rtol=0, atol=1e-11)


def test_module_keeps_the_reference_parameters_and_falls_back_on_cpu():
    import quan_ultralytics_b200 as Q
    m = Q.QER(64, 15, 1).double()
    assert sorted(k for k, _ in m.named_parameters(remove_duplicate=False)) == ["bias", "output_proj.bias", "output_proj.weight"]
    assert m.bias is m.output_proj.bias
    x = torch.randn(2, 16, 5, 6, 4, dtype=torch.float64)
    assert not m.fused_ok(x)                                    # CPU tensor: the view + library conv path, still the reference's result
    y = m(x)
    want = O.qer_fwd(x.numpy(), m.output_proj.weight.detach().numpy(), m.output_proj.bias.detach().numpy())
    np.testing.assert_allclose(y.detach().numpy(), want, rtol=0, atol=1e-12)


def _module(case, dtype):
    import quan_ultralytics_b200 as Q
    w, b = _g(case, "w"), _g(case, "b")
    m = Q.QER(w.shape[1], w.shape[0], 1).to(DEV)
    with torch.no_grad():
        m.output_proj.weight.copy_(torch.from_numpy(w))
        m.output_proj.bias.copy_(torch.from_numpy(b))
    x = torch.from_numpy(_g(case, "x")).to(DEV, dtype).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    return m, x


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_fp32_kernels_match_golden(case):
    m, x = _module(case, torch.float32)
    assert m.fused_ok(x)
    y = m(x)
    assert y.shape == _g(case, "y").shape and rel(y.detach().cpu(), _g(case, "y")) <= 1e-5
    y.backward(torch.from_numpy(_g(case, "dy")).to(DEV, torch.float32))
    assert rel(x.grad.cpu(), _g(case, "dx")) <= 1e-5
    assert rel(m.output_proj.weight.grad.cpu(), _g(case, "dw")) <= 1e-5
    assert rel(m.output_proj.bias.grad.cpu(), _g(case, "db")) <= 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_bf16_kernels_match_oracle_on_rounded_operands(case):
    m, x = _module(case, torch.bfloat16)
    assert m.fused_ok(x)
    xr = x.detach().double().cpu().numpy()
    wr = m.output_proj.weight.detach().bfloat16().double().cpu().numpy()           # the kernel rounds the staged weight to bf16
    b = _g(case, "b").astype(np.float32).astype(np.float64)
    y = m(x)
    assert y.dtype == torch.bfloat16 and rel(y.detach().double().cpu(), O.qer_fwd(xr, wr, b)) <= 1e-2
    dy = torch.from_numpy(_g(case, "dy")).to(DEV, torch.bfloat16)
    y.backward(dy)
    dx, dw, db = O.qer_bwd(dy.double().cpu().numpy(), xr, wr)
    assert rel(x.grad.double().cpu(), dx) <= 1e-2
    assert m.output_proj.weight.grad.dtype == torch.float32
    assert rel(m.output_proj.weight.grad.double().cpu(), dw) <= 2e-3              # fp32 accumulation of exact bf16 products
    assert rel(m.output_proj.bias.grad.double().cpu(), db) <= 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_qer_cat_writes_and_reads_the_concatenated_head_tensor(dtype):
    """functional.qer_cat == torch.cat((qer_a(xa), qer_b(xb)), 1) (head.py:143): forward columns in place, backward from the channel
    slices of the concatenated gradient (row pitch 79 elements: unaligned rows)."""
    import quan_ultralytics_b200 as Q
    from quan_ultralytics_b200 import functional as QF
    torch.manual_seed(4)
    B, C, H, W = 2, 16, 19, 21
    qa, qb = Q.QER(4 * C, 64, 1).to(DEV), Q.QER(4 * C, 15, 1).to(DEV)
    xa = torch.randn(B, C, H, W, 4, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    xb = torch.randn(B, C, H, W, 4, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    out = QF.qer_cat(xa, qa.output_proj.weight, qa.output_proj.bias, xb, qb.output_proj.weight, qb.output_proj.bias)
    assert out.shape == (B, 79, H, W)
    dy = torch.randn(B, 79, H, W, device=DEV).to(dtype)
    out.backward(dy, retain_graph=True)                        # dense NCHW gradient: made channels-last, rows 79 apart (unaligned)
    first = [t.grad.clone() for t in (xa, xb, qa.output_proj.weight, qb.output_proj.weight, qa.output_proj.bias, qb.output_proj.bias)]
    for t in (xa, xb, qa.output_proj.weight, qb.output_proj.weight, qa.output_proj.bias, qb.output_proj.bias):
        t.grad = None
    assert out.stride(3) == 80                                 # the fused head tensor pads its rows to 16 bytes ...
    pad = torch.zeros(B, H, W, 80, device=DEV, dtype=dtype)
    pad[..., :79] = dy.permute(0, 2, 3, 1)
    out.backward(pad[..., :79].permute(0, 3, 1, 2))             # ... and the loss hands its gradient back in the same rows (vector loads)
    for a, t in zip(first, (xa, xb, qa.output_proj.weight, qb.output_proj.weight, qa.output_proj.bias, qb.output_proj.bias)):
        assert torch.equal(a, t.grad)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    rnd = (lambda t: t.detach().double().cpu().numpy()) if dtype == torch.float32 else (lambda t: t.detach().bfloat16().double().cpu().numpy())
    got_w = {}
    for q, x, sl in ((qa, xa, slice(0, 64)), (qb, xb, slice(64, 79))):
        w, b = rnd(q.output_proj.weight), q.output_proj.bias.detach().double().cpu().numpy()
        xr = x.detach().double().cpu().numpy()
        assert rel(out[:, sl].detach().double().cpu(), O.qer_fwd(xr, w, b)) <= tol
        dx, dw, db = O.qer_bwd(dy[:, sl].double().cpu().numpy(), xr, w)
        assert rel(x.grad.double().cpu(), dx) <= tol
        assert rel(q.output_proj.weight.grad.double().cpu(), dw) <= max(tol / 5, 1e-5)
        assert rel(q.output_proj.bias.grad.double().cpu(), db) <= max(tol / 5, 1e-5)


@pytest.mark.gpu
def test_bf16_kernels_at_the_head_size_against_the_library_path():
    """BASELINE configs[2] P3 level (16 x 128 x 128 pixels, 64 -> 64): the kernel against torch's own conv on the re-ordered weight
    (the path QER.forward falls back to), bf16; plus linearity, a size-independent property: qer(a x) - bias = a (qer(x) - bias)."""
    import quan_ultralytics_b200 as Q
    torch.manual_seed(9)
    B, C, H, W, N = 16, 16, 128, 128, 64
    m = Q.QER(4 * C, N, 1).to(DEV)
    x = torch.randn(B, C, H, W, 4, device=DEV).bfloat16().contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    y = m(x)
    conv = m.output_proj
    w = conv.weight.view(N, C, 4, 1, 1).transpose(1, 2).reshape(N, 4 * C, 1, 1)
    xr = x.detach().permute(0, 4, 1, 2, 3).reshape(B, 4 * C, H, W).float().requires_grad_(True)
    ref = torch.nn.functional.conv2d(xr, w.bfloat16().float(), conv.bias)
    assert rel(y.detach().float().cpu(), ref.detach().cpu()) <= 1e-2
    dy = torch.randn_like(ref)
    gw_ref, gx_ref = torch.autograd.grad(ref, (conv.weight, xr), dy)
    y.backward(dy.bfloat16())
    gx = x.grad.permute(0, 4, 1, 2, 3).reshape(B, 4 * C, H, W)
    assert rel(gx.float().cpu(), gx_ref.cpu()) <= 1e-2
    assert rel(conv.weight.grad.cpu(), gw_ref.cpu()) <= 1e-2
    with torch.no_grad():
        y2 = m((x * 2).detach())
        lin = (y2.float() - conv.bias.view(1, -1, 1, 1)) - 2 * (y.float() - conv.bias.view(1, -1, 1, 1))
    assert float(lin.abs().max()) <= 2e-2 * float(y.float().abs().max())
