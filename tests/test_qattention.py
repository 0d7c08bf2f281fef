"""QAttention core: oracle/quan_oracle.py qattention_fwd/bwd pinned to golden vectors recorded from the real
reference module (tests/golden/make_qattn_golden.py, block.py:1485-1546), and quan_qattention_fwd/bwd on the GPU against both."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import quan_oracle as O

G = np.load(Path(__file__).resolve().parent / "golden" / "qattn.npz")


def _case(tag):
    heads, K, V, B, H, W = (int(v) for v in G[f"{tag}_meta"])
    return heads, K, V, float(G[f"{tag}_scale"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_matches_reference_module(tag):
    heads, K, V, scale = _case(tag)
    qkv = G[f"{tag}_qkv"]
    o, _ = O.qattention_fwd(qkv, heads, K, V, scale)
    # finish the module with the oracle's own convs: y = proj(o + pe(o))   (block.py:1543-1544)
    pe = [G[f"{tag}_pe_w{c}"] for c in "rijk"]
    proj = [G[f"{tag}_proj_w{c}"] for c in "rijk"]
    C = o.shape[1]
    y = O.qconv2d_fwd(o + O.qconv2d_fwd(o, pe, None, 1, 1, 1, C, O.M_A), proj, None, 1, 0, 1, 1, O.M_A)
    np.testing.assert_allclose(y, G[f"{tag}_y"], rtol=1e-10, atol=1e-12)
    # backward of the core: feed the gradient that reaches o in the reference (via the oracle's conv backward)
    d1, _, _ = O.qconv2d_bwd(G[f"{tag}_dy"], o + O.qconv2d_fwd(o, pe, None, 1, 1, 1, C, O.M_A), proj, 1, 0, 1, 1, O.M_A)
    d2, _, _ = O.qconv2d_bwd(d1, o, pe, 1, 1, 1, C, O.M_A)
    dqkv = O.qattention_bwd(d1 + d2, qkv, heads, K, V, scale)
    np.testing.assert_allclose(dqkv, G[f"{tag}_dqkv"], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 1e-2)])
def test_kernel_matches_oracle_on_golden_inputs(tag, dtype, tol):
    from quan_ultralytics_b200 import functional as QF
    heads, K, V, scale = _case(tag)
    qkv = torch.from_numpy(G[f"{tag}_qkv"]).to("cuda", dtype)
    qkv64 = qkv.double().cpu().numpy()                       # the oracle sees the same (rounded) inputs
    x = qkv.contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
    o = QF.qattention(x, heads, K, V, scale)
    d_o = torch.randn(o.shape, generator=torch.Generator().manual_seed(3)).to("cuda", dtype)
    o.backward(d_o)
    torch.cuda.synchronize()
    o_ref, _ = O.qattention_fwd(qkv64, heads, K, V, scale)
    g_ref = O.qattention_bwd(d_o.double().cpu().numpy(), qkv64, heads, K, V, scale)
    rel = lambda a, b: float(np.abs(a.detach().double().cpu().numpy() - b).max() / np.abs(b).max())
    assert rel(o, o_ref) <= tol, rel(o, o_ref)
    assert rel(x.grad, g_ref) <= 2 * tol, rel(x.grad, g_ref)


@pytest.mark.gpu
def test_kernel_at_model_size_vs_torch_restatement():
    """QUAN-YOLO11n P5 at 1024^2: N = 1024 tokens, 8 heads of (2, 4); B = 2 — against the reference's own op sequence run in fp32 on
    the device (split / reshape / matmul / softmax / matmul, block.py:1520-1540); also N not a multiple of the block (30 x 30)."""
    from quan_ultralytics_b200 import functional as QF
    for B, H, W in [(2, 32, 32), (1, 30, 30), (1, 40, 48)]:
        heads, K, V = 8, 2, 4
        torch.manual_seed(0)
        qkv = torch.randn(B, heads * (2 * K + V), H, W, 4, device="cuda")
        x = qkv.clone().contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
        o = QF.qattention(x, heads, K, V, K ** -0.5)
        d_o = torch.randn_like(o)
        o.backward(d_o)
        xr = qkv.clone().requires_grad_(True)
        N = H * W
        q, k, v = torch.split(xr, [heads * K, heads * K, heads * V], dim=1)
        q = q.reshape(B, heads, K, N, 4).permute(0, 1, 4, 3, 2)
        k = k.reshape(B, heads, K, N, 4).permute(0, 1, 4, 2, 3)
        v = v.reshape(B, heads, V, N, 4).permute(0, 1, 4, 3, 2)
        attn = (torch.matmul(q, k) * K ** -0.5).softmax(dim=-1)
        orf = torch.matmul(attn, v).permute(0, 1, 4, 3, 2).reshape(B, heads * V, H, W, 4)
        orf.backward(d_o)
        rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
        assert rel(o, orf) <= 1e-4 and rel(x.grad, xr.grad) <= 1e-4, (B, H, W, rel(o, orf), rel(x.grad, xr.grad))
