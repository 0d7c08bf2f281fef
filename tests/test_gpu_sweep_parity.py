"""Independent parity of the tcgen05 engine at BASELINE sweep sizes (configs[1]: C_q = 128 / 256 / 512, 3x3, stride 1 and 2 — halo mode,
persistent multi-unit CTAs, stride-2 parity-class dgrad, K = 4608 in tf32) — NOT against the library's own direct engine but against
(a) oracle/torch_port.py (the call-by-call restatement of the reference's PyTorch path, conv.py:472-499) run on the device in fp32 with
TF32 disabled, and (b) the reference's own CUDA extension (oracle/_ref/quaternion_ops.so, M_B) where it is present.
Tolerances are BASELINE.json's: 1e-3 (fp32 storage / tf32 MMA), 1e-2 (bf16); the measured margins are printed."""
import pytest
import torch

from oracle import torch_port as TP
from quan_ultralytics_b200 import ops

pytestmark = pytest.mark.gpu
L = ops.LAYOUT_BHWQC

CASES = [  # name, dtype, N, C, H, stride, mix
    ("tf32_c256_s1", "f32", 8, 256, 32, 1, "A"),
    ("tf32_c512_s1_K4608", "f32", 8, 512, 16, 1, "A"),
    ("tf32_c512_s2", "f32", 6, 512, 16, 2, "B"),
    ("tf32_c128_s2", "f32", 4, 128, 64, 2, "A"),
    ("bf16_c256_s1", "bf16", 8, 256, 32, 1, "A"),
    ("bf16_c512_s1", "bf16", 8, 512, 16, 1, "B"),
    ("bf16_c256_s2", "bf16", 8, 256, 32, 2, "A"),
    ("bf16_c128_s1_persist", "bf16", 40, 128, 64, 1, "A"),
]


@pytest.fixture()
def exact_fp32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tcgen05_engine_vs_reference_pytorch_path_at_sweep_sizes(exact_fp32, case):
    name, dt, N, C, H, s, mix = case
    dtype = torch.bfloat16 if dt == "bf16" else torch.float32
    tol = 1e-2 if dt == "bf16" else 1e-3
    torch.manual_seed(1234)
    x = torch.randn(N, C, H, H, 4, device="cuda").to(dtype)
    w = [torch.randn(C, C, 3, 3, device="cuda") / (C * 9) ** 0.5 for _ in range(4)]
    args = ((s, s), (1, 1), (1, 1), 1)
    assert all(ops.qconv2d_pick_algo(x.shape, w[0].shape, *args, dtype, L, ps) == ops.ALGO_TCGEN05 for ps in range(3))
    xl = x.contiguous(memory_format=torch.channels_last_3d)
    y = ops.qconv2d_fwd(xl, w, None, *args, ops.MIX[mix], ops.ALGO_TCGEN05, L)
    dy = torch.randn(y.shape, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last_3d)
    dx, dw, _ = ops.qconv2d_bwd(dy, xl, w, *args, ops.MIX[mix], True, True, False)
    # oracle: the reference's op sequence in fp32 (exact convolutions), on the same (rounded) inputs; bf16 weights as the engine multiplies them
    xr = x.float().requires_grad_(True)
    wr = [(t.to(dtype).float() if dt == "bf16" else t.clone()).requires_grad_(True) for t in w]
    yr = TP.qconv2d(xr, *wr, None, *args, mix)
    gx, *gw = torch.autograd.grad(yr, [xr, *wr], dy.float())
    e = {"y": _rel(y, yr), "dx": _rel(dx, gx), "dw": max(_rel(a, b) for a, b in zip(dw, gw))}
    print(f"\n[{name}] vs fp32 PyTorch path: " + " ".join(f"{k}={v:.2e} ({100 * v / tol:.0f}% of {tol:g})" for k, v in e.items()))
    assert e["y"] <= tol and e["dx"] <= tol and e["dw"] <= tol, e


@pytest.mark.parametrize("C,H,s", [(256, 32, 1), (512, 16, 2)])
def test_tcgen05_engine_vs_reference_cuda_extension_at_sweep_sizes(C, H, s):
    """The reference's own kernels (quaternion_ops.cu, fp32 CUDA cores, M_B) on 2 images at the sweep shapes."""
    from oracle import build_ref_ext
    ref = build_ref_ext.load()
    if ref is None:
        pytest.skip("oracle/_ref/quaternion_ops.so not built (needs the reference checkout)")
    torch.manual_seed(5)
    N = 2
    x = torch.randn(N, C, H, H, 4, device="cuda")
    w = [torch.randn(C, C, 3, 3, device="cuda") / (C * 9) ** 0.5 for _ in range(4)]
    args = ([s, s], [1, 1], [1, 1], 1)
    y_ref = ref.qconv_forward(x, *w, None, None, None, None, *args)
    dy = torch.randn_like(y_ref)
    g_ref = ref.qconv_backward(dy, x, *w, False, *args)
    xl, dyl = (t.contiguous(memory_format=torch.channels_last_3d) for t in (x, dy))
    targs = ((s, s), (1, 1), (1, 1), 1, ops.M_B)
    y = ops.qconv2d_fwd(xl, w, None, *targs, ops.ALGO_TCGEN05, L)
    dx, dw, _ = ops.qconv2d_bwd(dyl, xl, w, *targs, True, True, False)
    e = {"y": _rel(y, y_ref), "dx": _rel(dx, g_ref[0]), "dw": max(_rel(a, b) for a, b in zip(dw, g_ref[1:5]))}
    print(f"\n[C={C} {H}x{H} s{s}] tcgen05 (tf32) vs the reference CUDA extension: " + " ".join(f"{k}={v:.2e}" for k, v in e.items()))
    assert max(e.values()) <= 1e-3, e
