"""GPU parity of the tcgen05/TMEM/TMA engine (qconv_tc.cu) through the C ABI: against the golden-validated direct
engine on identical device tensors, and against the CPU oracle at sizes it finishes in seconds."""
import numpy as np
import pytest
import torch

from quan_ultralytics_b200 import ops
from oracle import quan_oracle as O
from tests.tc_probe import CASES

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
L = ops.LAYOUT_BHWQC


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make(case, seed=0):
    name, dt, B, Ci, Co, H, W, k, s, p, d, bias, mix = case
    dtype = torch.bfloat16 if dt == "bf16" else torch.float32
    torch.manual_seed(seed)
    x = torch.randn(B, Ci, H, W, 4, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    w = [torch.randn(Co, Ci, k, k, device=DEV) / (Ci * k * k) ** 0.5 for _ in range(4)]
    b = torch.randn(Co, device=DEV) if bias else None
    return dtype, x, w, b, ((s, s), (p, p), (d, d), 1, ops.MIX[mix])


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_tc_matches_direct_engine(case):
    dtype, x, w, b, args = make(case)
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-3
    picks = [ops.qconv2d_pick_algo(x.shape, w[0].shape, *args[:4], dtype, L, ps) for ps in range(3)]
    assert picks[0] == ops.ALGO_TCGEN05, "every probe case is meant to qualify for the tensor-core forward"
    y_ref = ops.qconv2d_fwd(x, w, b, *args, ops.ALGO_DIRECT, L)
    y = ops.qconv2d_fwd(x, w, b, *args, ops.ALGO_TCGEN05, L)
    assert rel(y, y_ref) <= tol
    dy = torch.randn_like(y_ref)
    dx_ref, dw_ref, db_ref = ops.qconv2d_bwd(dy, x, w, *args, True, True, b is not None, ops.ALGO_DIRECT)
    dx, dw, db = ops.qconv2d_bwd(dy, x, w, *args, True, True, b is not None, ops.ALGO_AUTO)
    assert rel(dx, dx_ref) <= 2 * tol
    for a, r in zip(dw, dw_ref):
        assert rel(a, r) <= 2 * tol
    if b is not None:
        assert rel(db, db_ref) <= 1e-4


@pytest.mark.parametrize("dt", ["bf16", "f32"])
def test_tc_vs_oracle_fp64(dt):
    """Independent of the direct engine: fp64 numpy oracle on the same (bf16-rounded) inputs."""
    case = ("oracle", dt, 2, 64, 64, 12, 10, 3, 1, 1, 1, True, "A")
    dtype, x, w, b, args = make(case, seed=3)
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-3
    xn = x.double().cpu().numpy()
    wn = [t.double().cpu().numpy() for t in w]
    if dtype == torch.bfloat16:          # the engine multiplies bf16 copies of the master weights
        wn = [t.to(torch.bfloat16).double().cpu().numpy() for t in w]
    y_ref = O.qconv2d_fwd(xn, wn, b.double().cpu().numpy(), 1, 1, 1, 1, O.M_A)
    y = ops.qconv2d_fwd(x, w, b, *args, ops.ALGO_TCGEN05, L)
    assert float(np.max(np.abs(y.double().cpu().numpy() - y_ref)) / np.max(np.abs(y_ref))) <= tol
    dy = torch.randn_like(y)
    dx_ref, dw_ref, db_ref = O.qconv2d_bwd(dy.double().cpu().numpy(), xn, wn, 1, 1, 1, 1, O.M_A, has_bias=True)
    dx, dw, db = ops.qconv2d_bwd(dy, x, w, *args, True, True, True, ops.ALGO_TCGEN05)
    nerr = lambda a, r: float(np.max(np.abs(a.double().cpu().numpy() - r)) / np.max(np.abs(r)))
    assert nerr(dx, dx_ref) <= 2 * tol
    for q in range(4):
        assert nerr(dw[q], dw_ref[q]) <= 2 * tol
    assert nerr(db, db_ref) <= 2 * tol


def test_tc_is_what_auto_picks_for_the_sweep_shapes():
    for C in (64, 128, 256, 512):
        for dtype in (torch.bfloat16, torch.float32):
            for s in (1, 2):
                picks = [ops.qconv2d_pick_algo((16, C, 32, 32, 4), (C, C, 3, 3), (s, s), (1, 1), (1, 1), 1, dtype, L, ps)
                         for ps in range(3)]
                assert picks[0] == ops.ALGO_TCGEN05 and picks[2] == ops.ALGO_TCGEN05
                assert picks[1] == ops.ALGO_TCGEN05   # stride 2: parity-class dgrad
    # narrow layers take the dense Hamilton form of the tensor-core engine
    assert ops.qconv2d_pick_algo((16, 4, 32, 32, 4), (8, 4, 3, 3), (1, 1), (1, 1), (1, 1), 1, torch.bfloat16, L, 0) \
        == ops.ALGO_TCGEN05
    # reference layout / grouped / 1-2 channel convs stay on the direct engine
    assert ops.qconv2d_pick_algo((16, 64, 32, 32, 4), (64, 64, 3, 3), (1, 1), (1, 1), (1, 1), 1, torch.bfloat16,
                                 ops.LAYOUT_BCHWQ, 0) == ops.ALGO_DIRECT
    # depthwise layers of up to 64 quaternion channels: dense form with a block-diagonal weight on the tensor cores; wider ones on the
    # streaming CUDA-core engine
    assert ops.qconv2d_pick_algo((16, 64, 32, 32, 4), (64, 1, 3, 3), (1, 1), (1, 1), (1, 1), 64, torch.bfloat16, L, 0) \
        == ops.ALGO_TCGEN05
    assert ops.qconv2d_pick_algo((16, 128, 32, 32, 4), (128, 1, 3, 3), (1, 1), (1, 1), (1, 1), 128, torch.bfloat16, L, 0) \
        == ops.ALGO_DEPTHWISE
    assert ops.qconv2d_pick_algo((16, 64, 32, 32, 4), (64, 2, 3, 3), (1, 1), (1, 1), (1, 1), 32, torch.bfloat16, L, 0) \
        == ops.ALGO_DIRECT
    assert ops.qconv2d_pick_algo((16, 2, 32, 32, 4), (4, 2, 3, 3), (1, 1), (1, 1), (1, 1), 1, torch.bfloat16, L, 0) \
        == ops.ALGO_SMALLC
    assert ops.qconv2d_pick_algo((16, 3, 32, 32, 4), (5, 3, 3, 3), (1, 1), (1, 1), (1, 1), 1, torch.bfloat16, L, 0) \
        == ops.ALGO_DIRECT


# name: (dtype, B, C_i, C_o, H, W, k, s, p)
SPARSE_PARITY_CASES = [
    ("qresnet_shortcut_dense", "bf16", 5, 16, 32, 28, 28, 1, 2, 0),
    ("qresnet_shortcut_sep", "bf16", 3, 64, 128, 14, 14, 1, 2, 0),
    ("k1_s2_odd_f32", "f32", 2, 64, 64, 15, 13, 1, 2, 0),
    ("k1_s2_dense_ragged", "bf16", 2, 8, 16, 21, 17, 1, 2, 0),
]


@pytest.mark.parametrize("case", SPARSE_PARITY_CASES, ids=[c[0] for c in SPARSE_PARITY_CASES])
def test_strided_dgrad_with_empty_parity_classes(case):
    """Strided dgrad whose filter misses some output parities (1x1 stride 2, the Q-ResNet-34 shortcut convs,
    classification/models/quaternion_models.py): those classes are dropped, dX is cleared first.  Against the direct engine."""
    name, dt, B, ci, co, H, W, k, s, p = case
    dtype = torch.bfloat16 if dt == "bf16" else torch.float32
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-3
    torch.manual_seed(3)
    x = torch.randn(B, ci, H, W, 4, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    w = [torch.randn(co, ci, k, k, device=DEV) / (ci * k * k) ** 0.5 for _ in range(4)]
    args = ((s, s), (p, p), (1, 1), 1, ops.M_B)
    assert ops.qconv2d_pick_algo(x.shape, w[0].shape, *args[:4], dtype, L, 1) == ops.ALGO_TCGEN05
    y = ops.qconv2d_fwd(x, w, None, *args, ops.ALGO_AUTO, L)
    dy = torch.randn_like(y)
    dx_ref, dw_ref, _ = ops.qconv2d_bwd(dy, x, w, *args, True, True, False, ops.ALGO_DIRECT)
    dx, dw, _ = ops.qconv2d_bwd(dy, x, w, *args, True, True, False, ops.ALGO_AUTO)
    assert rel(dx, dx_ref) <= 2 * tol
    # rows / columns no tap reaches are exactly zero
    if k == 1:
        assert float(dx[:, :, 1::2].abs().max()) == 0.0 and float(dx[:, :, :, 1::2].abs().max()) == 0.0
    for a, r in zip(dw, dw_ref):
        assert rel(a, r) <= 2 * tol


def test_tc_linearity_at_sweep_size():
    """BASELINE-size property (no oracle): conv(a + 2b) == conv(a) + 2 conv(b) through the tensor-core path."""
    torch.manual_seed(0)
    C = 128
    w = [torch.randn(C, C, 3, 3, device=DEV) * 0.03 for _ in range(4)]
    a = torch.randn(16, C, 64, 64, 4, device=DEV).contiguous(memory_format=torch.channels_last_3d)
    b = torch.randn_like(a)
    args = ((1, 1), (1, 1), (1, 1), 1, ops.M_B)
    f = lambda t: ops.qconv2d_fwd(t, w, None, *args, ops.ALGO_TCGEN05, L)
    assert rel(f(a + 2 * b), f(a) + 2 * f(b)) <= 2e-3


# name: (dtype, B, C, H, W, k, s, p, d, bias, mix)
DW_CASES = [
    ("bf16_c16", "bf16", 3, 16, 20, 24, 3, 1, 1, 1, False, "A"),
    ("bf16_c64_bias", "bf16", 2, 64, 16, 16, 3, 1, 1, 1, True, "B"),
    ("bf16_c12_s2", "bf16", 2, 12, 17, 15, 3, 2, 1, 1, False, "A"),
    ("f32_c8_dil2", "f32", 2, 8, 14, 14, 3, 1, 2, 2, True, "A"),
    ("f32_c6_k5", "f32", 2, 6, 12, 12, 5, 1, 2, 1, False, "B"),
    ("bf16_c32_yolo", "bf16", 4, 32, 64, 64, 3, 1, 1, 1, False, "A"),
]


@pytest.mark.parametrize("case", DW_CASES, ids=[c[0] for c in DW_CASES])
def test_depthwise_engine_matches_direct_engine(case):
    """The streaming depthwise kernels (qconv_dw.cu; DWConv, conv.py:918-923) against the golden-validated generic engine."""
    name, dt, B, Cq, H, W, k, s, p, d, bias, mix = case
    dtype = torch.bfloat16 if dt == "bf16" else torch.float32
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-5      # both engines are true fp32 FMA: only summation order differs
    torch.manual_seed(5)
    x = torch.randn(B, Cq, H, W, 4, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    w = [torch.randn(Cq, 1, k, k, device=DEV) / k for _ in range(4)]
    b = torch.randn(Cq, device=DEV) if bias else None
    args = ((s, s), (p, p), (d, d), Cq, ops.MIX[mix])
    y_ref = ops.qconv2d_fwd(x, w, b, *args, ops.ALGO_DIRECT, L)
    y = ops.qconv2d_fwd(x, w, b, *args, ops.ALGO_DEPTHWISE, L)
    assert rel(y, y_ref) <= tol
    dy = torch.randn_like(y_ref)
    dx_ref, dw_ref, db_ref = ops.qconv2d_bwd(dy, x, w, *args, True, True, bias, ops.ALGO_DIRECT)
    if k * k <= 9:
        dx, dw, db = ops.qconv2d_bwd(dy, x, w, *args, True, True, bias, ops.ALGO_DEPTHWISE)
    else:                                                # 5x5: the depthwise engine serves forward and dgrad, the generic engine wgrad
        dx, _, _ = ops.qconv2d_bwd(dy, x, w, *args, True, False, False, ops.ALGO_DEPTHWISE)
        _, dw, db = ops.qconv2d_bwd(dy, x, w, *args, False, True, bias, ops.ALGO_DIRECT)
    # the generic engine rounds G = M^T dY to bf16 before using it; the depthwise kernels keep it in fp32
    assert rel(dx, dx_ref) <= 2 * tol
    for a, r in zip(dw, dw_ref):
        assert rel(a, r) <= 2 * tol
    if bias:
        assert rel(db, db_ref) <= 1e-4


# name: (dtype, B, C, H, W, s, bias, mix) — depthwise 3x3 layers of the QUAN-YOLO11n head (16 / 32 / 64 channels) and ragged maps
DWTC_CASES = [("bf16_c16_128", "bf16", 2, 16, 128, 128, 1, False, "A"), ("bf16_c32_64", "bf16", 4, 32, 64, 64, 1, False, "A"),
              ("bf16_c64_32_bias", "bf16", 4, 64, 32, 32, 1, True, "B"), ("bf16_c16_ragged", "bf16", 3, 16, 21, 19, 1, False, "A"),
              ("f32_c16", "f32", 2, 16, 32, 32, 1, True, "A"), ("bf16_c32_s2", "bf16", 2, 32, 32, 32, 2, False, "A")]


@pytest.mark.parametrize("case", DWTC_CASES, ids=[c[0] for c in DWTC_CASES])
def test_depthwise_on_the_tensor_cores_matches_direct_engine(case):
    """DWConv (conv.py:918-923) of up to 64 quaternion channels as the dense tensor-core form with a block-diagonal packed weight
    (qconv_tc.cu depthwise_as_dense) against the golden-validated generic engine: forward, dgrad, and the wgrad whose fold keeps the
    block diagonal.  Tolerances are the engine's: 1e-2 bf16, 1e-3 tf32."""
    name, dt, B, Cq, H, W, s, bias, mix = case
    dtype = torch.bfloat16 if dt == "bf16" else torch.float32
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-3
    torch.manual_seed(6)
    x = torch.randn(B, Cq, H, W, 4, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    w = [torch.randn(Cq, 1, 3, 3, device=DEV) / 3 for _ in range(4)]
    b = torch.randn(Cq, device=DEV) if bias else None
    args = ((s, s), (1, 1), (1, 1), Cq, ops.MIX[mix])
    picks = [ops.qconv2d_pick_algo(x.shape, w[0].shape, *args[:4], dtype, L, ps) for ps in range(3)]
    assert picks[0] == ops.ALGO_TCGEN05 and (s != 1 or picks == [ops.ALGO_TCGEN05] * 3), picks
    y_ref = ops.qconv2d_fwd(x, w, b, *args, ops.ALGO_DIRECT, L)
    y = ops.qconv2d_fwd(x, w, b, *args, ops.ALGO_AUTO, L)
    assert rel(y, y_ref) <= tol
    dy = torch.randn_like(y_ref)
    dx_ref, dw_ref, db_ref = ops.qconv2d_bwd(dy, x, w, *args, True, True, bias, ops.ALGO_DIRECT)
    dx, dw, db = ops.qconv2d_bwd(dy, x, w, *args, True, True, bias, ops.ALGO_AUTO)
    assert dx.shape == x.shape and all(a.shape == (Cq, 1, 3, 3) for a in dw)
    assert rel(dx, dx_ref) <= 2 * tol
    for a, r in zip(dw, dw_ref):
        assert rel(a, r) <= 2 * tol
    if bias:
        assert rel(db, db_ref) <= 1e-3


@pytest.mark.parametrize("cfg", [("sep_bf16", torch.bfloat16, 64, 64, 1), ("sep_f32", torch.float32, 64, 128, 1),
                                 ("dense_bf16", torch.bfloat16, 16, 16, 1), ("dw_bf16", torch.bfloat16, 32, 32, 32),
                                 ("sep_bf16_s2", torch.bfloat16, 64, 64, 1)], ids=lambda c: c[0])
def test_fused_conv_block_matches_separate_nodes(cfg):
    """`Conv` as one autograd node (IQBN backward emits G = M^T dY for the conv backward) against the three separate
    nodes, same kernels otherwise: outputs identical, gradients equal up to the bf16 rounding of the intermediate."""
    import quan_ultralytics_b200 as Q
    name, dtype, ci, co, g = cfg
    s = 2 if name.endswith("s2") else 1
    torch.manual_seed(11)
    blk = (Q.DWConv(ci * 4, co * 4, 3, s) if g > 1 else Q.Conv(ci * 4, co * 4, 3, s)).to(DEV).train()
    x = torch.randn(4, ci, 24, 24, 4, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    outs = {}
    for fused in (True, False):
        blk.fuse_block = fused
        blk.zero_grad(set_to_none=True)
        blk.bn.running_mean.zero_(); blk.bn.running_var.fill_(1.0)
        xi = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
            y = blk(xi)
        torch.manual_seed(12)
        y.backward(torch.randn_like(y))
        outs[fused] = (y.detach(), xi.grad, blk.conv.weight_r.grad.clone(), blk.conv.weight_k.grad.clone(),
                       blk.bn.gamma.grad.clone(), blk.bn.beta.grad.clone(), blk.bn.running_var.clone())
    # fp32: G is the same value up to FMA grouping, then truncated to tf32 by the tensor core
    tol = 1e-2 if dtype == torch.bfloat16 else 1e-4
    # the fused node takes the batch statistics from the conv epilogue (fp32 accumulators, before y is rounded to bf16)
    assert rel(outs[True][0], outs[False][0]) <= tol and rel(outs[True][6], outs[False][6]) <= tol
    for a, b in zip(outs[True][1:6], outs[False][1:6]):
        assert rel(a, b) <= tol


@pytest.mark.parametrize("cfg", [("sep_bf16", "bf16", 9, 64, 128, 30, 26, 3, 1), ("sep_f32", "f32", 5, 64, 64, 17, 19, 3, 2),
                                 ("dense_bf16", "bf16", 40, 16, 32, 64, 64, 3, 1), ("dense_k1", "bf16", 7, 8, 16, 33, 31, 1, 1),
                                 ("sep_persist", "bf16", 70, 64, 64, 32, 32, 3, 1)], ids=lambda c: c[0])
def test_conv_epilogue_emits_iqbn_statistics(cfg):
    """quan_qconv2d_fwd_stats: per-CTA partial sums from the tensor-core epilogue, folded by quan_iqbn_finalize_partials,
    against the streaming statistics kernel on the stored output (ragged tiles, masked spare tiles, multi-unit CTAs)."""
    name, dt, B, ci, co, H, W, k, s_ = cfg
    dtype = torch.bfloat16 if dt == "bf16" else torch.float32
    torch.manual_seed(21)
    x = (torch.randn(B, ci, H, W, 4, device=DEV) + 0.3).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    w = [torch.randn(co, ci, k, k, device=DEV) / (ci * k * k) ** 0.5 for _ in range(4)]
    args = ((s_, s_), (k // 2, k // 2), (1, 1), 1, ops.M_A)
    y, nparts = ops.qconv2d_fwd(x, w, None, *args, ops.ALGO_AUTO, L, with_stats=True)
    if name.startswith("sep"):
        # the separable (wide) form only emits statistics when forced (QUAN_TC_EPI_STATS=2): its epilogue has no slack
        import os
        if os.environ.get("QUAN_TC_EPI_STATS") != "2":
            assert nparts == 0
            pytest.skip("separable-form epilogue statistics are opt-in (QUAN_TC_EPI_STATS=2)")
    assert nparts > 0, "the dense form emits the partial sums"
    g, b = torch.rand(co, 4, device=DEV) + 0.5, torch.randn(co, 4, device=DEV)
    rm, rv = torch.zeros(co, 4, device=DEV), torch.ones(co, 4, device=DEV)
    cnt = float(y.shape[0] * y.shape[2] * y.shape[3])
    st = ops.iqbn_finalize_partials(nparts, cnt, co, g, b, 1e-5, 0.1, rm, rv)
    rm2, rv2 = torch.zeros(co, 4, device=DEV), torch.ones(co, 4, device=DEV)
    st_ref = ops.iqbn_train_stats(y, L, g, b, 1e-5, 0.1, rm2, rv2)
    tol = 4e-3 if dtype == torch.bfloat16 else 1e-5      # bf16: the reference pass sees y after rounding to bf16
    n = 4 * co
    for lo, hi in ((0, n), (n, 2 * n), (2 * n, 3 * n), (3 * n, 4 * n), (4 * n, 5 * n)):   # mean | var | rstd | scaleT | shiftT
        assert rel(st[lo:hi], st_ref[lo:hi]) <= tol
    assert rel(rm, rm2) <= tol and rel(rv, rv2) <= tol


# name: (dtype, B, Ci, Co, H, W, k, s, p, d, bias, mix)
SMALL_CASES = [
    ("stem_1_4_s2", "bf16", 2, 1, 4, 64, 64, 3, 2, 1, 1, False, "A"),
    ("yolo_4_2", "bf16", 2, 4, 2, 33, 31, 3, 1, 1, 1, False, "A"),
    ("yolo_2_4", "bf16", 2, 2, 4, 32, 32, 3, 1, 1, 1, True, "B"),
    ("f32_8_2_k5", "f32", 2, 8, 2, 20, 20, 5, 1, 2, 1, True, "A"),
    ("f32_1_8_dil2_s2", "f32", 3, 1, 8, 21, 23, 3, 2, 2, 2, False, "B"),
    ("f32_2_2_k1", "f32", 2, 2, 2, 16, 16, 1, 1, 0, 1, False, "A"),
    # wide outputs behind 1..2 input channels run forward / wgrad in chunks of 8 output channels (Q-ResNet-34 stem)
    ("qresnet_stem_1_16_k7", "bf16", 2, 1, 16, 40, 36, 7, 2, 3, 1, True, "B"),
    ("chunk_2_24_bf16", "bf16", 2, 2, 24, 19, 17, 3, 1, 1, 1, False, "A"),
]


@pytest.mark.parametrize("case", SMALL_CASES, ids=[c[0] for c in SMALL_CASES])
def test_small_channel_engine_matches_direct_engine(case):
    """One-thread-per-pixel kernels for 1..8 quaternion channels (qconv_small.cu; the QUAN-YOLO11n stem) against the
    golden-validated generic engine."""
    name, dt, B, Ci, Co, H, W, k, s, p, d, bias, mix = case
    dtype = torch.bfloat16 if dt == "bf16" else torch.float32
    tol = 1e-2 if dtype == torch.bfloat16 else 2e-5
    torch.manual_seed(9)
    x = torch.randn(B, Ci, H, W, 4, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    w = [torch.randn(Co, Ci, k, k, device=DEV) / (Ci * k * k) ** 0.5 for _ in range(4)]
    b = torch.randn(Co, device=DEV) if bias else None
    args = ((s, s), (p, p), (d, d), 1, ops.MIX[mix])
    picks = [ops.qconv2d_pick_algo(x.shape, w[0].shape, *args[:4], dtype, L, ps) for ps in range(3)]
    # the tensor-core dense form has priority where its row / N granularity allows (e.g. fp32 dgrad with 4*C_o*4 B = 32 B rows)
    chunked = Co > 8
    assert picks[0] == ops.ALGO_SMALLC and all(pk in (ops.ALGO_SMALLC, ops.ALGO_TCGEN05) for pk in picks[::2])
    assert picks[1] in ((ops.ALGO_DIRECT, ops.ALGO_TCGEN05) if chunked else (ops.ALGO_SMALLC, ops.ALGO_TCGEN05))
    if chunked:
        assert picks[2] == ops.ALGO_SMALLC
    if dtype == torch.float32 and ops.ALGO_TCGEN05 in picks:
        tol = 1e-3
    y_ref = ops.qconv2d_fwd(x, w, b, *args, ops.ALGO_DIRECT, L)
    y = ops.qconv2d_fwd(x, w, b, *args, ops.ALGO_AUTO, L)
    assert rel(y, y_ref) <= tol
    dy = torch.randn_like(y_ref)
    dx_ref, dw_ref, db_ref = ops.qconv2d_bwd(dy, x, w, *args, True, True, bias, ops.ALGO_DIRECT)
    dx, dw, db = ops.qconv2d_bwd(dy, x, w, *args, True, True, bias, ops.ALGO_AUTO)
    assert rel(dx, dx_ref) <= 2 * tol
    for a, r in zip(dw, dw_ref):
        assert rel(a, r) <= 2 * tol
    if bias:
        assert rel(db, db_ref) <= 1e-4


@pytest.mark.parametrize("cfg", [("bf16_c160", "bf16", 3, 160, 9, 11), ("f32_c96", "f32", 2, 96, 10, 7),
                                 ("bf16_c24", "bf16", 5, 24, 17, 13), ("bf16_c512", "bf16", 2, 512, 6, 5)], ids=lambda c: c[0])
def test_iqbn_streaming_kernels_vs_oracle_wide_rows(cfg):
    """TMA-fed IQBN reductions / backward apply (rows of 4*C_q > 256 parameters, ragged last tile, fused G = M^T dx) against
    the fp64 oracle on the same (rounded) inputs."""
    import quan_ultralytics_b200 as Q
    name, dt, B, C_, H, W = cfg
    dtype = torch.bfloat16 if dt == "bf16" else torch.float32
    tol = TOL = 1e-2 if dtype == torch.bfloat16 else 1e-4
    rng = np.random.default_rng(7)
    x = rng.normal(size=(B, C_, H, W, 4)) * 1.5 + 0.4
    dy = rng.normal(size=(B, C_, H, W, 4))
    gamma = rng.uniform(0.5, 1.5, size=(C_, 4))
    beta = rng.normal(size=(C_, 4))
    xt = torch.from_numpy(x).to(DEV, dtype).contiguous(memory_format=torch.channels_last_3d)
    dyt = torch.from_numpy(dy).to(DEV, dtype).contiguous(memory_format=torch.channels_last_3d)
    xr, dyr = xt.double().cpu().numpy(), dyt.double().cpu().numpy()      # what the kernels actually read
    g32 = torch.from_numpy(gamma).to(DEV, torch.float32)
    b32 = torch.from_numpy(beta).to(DEV, torch.float32)
    cnt = float(B * H * W)
    stats = ops.iqbn_train_stats(xt, L, g32, b32, 1e-5, 0.1, None, None)
    y = ops.iqbn_apply_fwd(xt, L, stats, g32, b32, Q.ACT_SILU)
    y_ref, _, _, _ = O.iqbn_train_fwd(xr, gamma, beta, act=True)
    assert float(np.max(np.abs(y.double().cpu().numpy() - y_ref)) / np.max(np.abs(y_ref))) <= tol
    sums = ops.iqbn_bwd_reduce(dyt, xt, L, stats, g32, b32, Q.ACT_SILU, cnt)
    dx, dg, db = ops.iqbn_bwd_apply(dyt, xt, L, stats, g32, b32, Q.ACT_SILU, sums, cnt)
    dx_ref, dg_ref, db_ref = O.iqbn_train_bwd(dyr, xr, gamma, beta, act=True)
    nerr = lambda a, r: float(np.max(np.abs(a.double().cpu().numpy() - r)) / np.max(np.abs(r)))
    assert nerr(dx, dx_ref) <= 2 * tol and nerr(dg, dg_ref) <= 2 * tol and nerr(db, db_ref) <= 2 * tol
    # fused G = M^T dx
    sums2 = ops.iqbn_bwd_reduce(dyt, xt, L, stats, g32, b32, Q.ACT_SILU, cnt)
    gmix, dg2, db2 = ops.iqbn_bwd_apply(dyt, xt, L, stats, g32, b32, Q.ACT_SILU, sums2, cnt, mix_t=ops._mix_t(ops.M_A))
    MT = np.array(ops.M_A).reshape(4, 4).T
    g_ref = np.einsum("qp,bchwp->bchwq", MT, dx_ref)
    assert nerr(gmix, g_ref) <= 2 * tol and nerr(dg2, dg_ref) <= 2 * tol and nerr(db2, db_ref) <= 2 * tol


@pytest.mark.parametrize("cfg", [("wide", 64, 64, 1), ("narrow", 16, 16, 1), ("depthwise", 32, 32, 32)], ids=lambda c: c[0])
def test_conv_block_step_captures_in_a_cuda_graph(cfg):
    """The whole Conv block step (weight packing, TMA descriptors, memsets, the dgrad || wgrad fork / join on the side
    stream) captures into a CUDA graph and replays to the same result as eager execution."""
    import quan_ultralytics_b200 as Q
    name, ci, co, g = cfg
    torch.manual_seed(3)
    blk = (Q.DWConv(ci * 4, co * 4, 3, 1) if g > 1 else Q.Conv(ci * 4, co * 4, 3, 1)).to(DEV).train()
    x = torch.randn(4, ci, 32, 32, 4, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    x.requires_grad_(True)
    dy = torch.randn(4, co, 32, 32, 4, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)

    def step():
        x.grad = None
        blk.zero_grad(set_to_none=True)
        y = blk(x)
        y.backward(dy)
        return y

    for _ in range(3):
        y_eager = step()
    ref = (y_eager.detach().clone(), x.grad.clone(), blk.conv.weight_r.grad.clone(), blk.bn.gamma.grad.clone())
    del y_eager      # a live autograd graph would pin its AccumulateGrad nodes to the default stream and break the capture
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y_static = step()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    got = (y_static.detach(), x.grad, blk.conv.weight_r.grad, blk.bn.gamma.grad)
    for a, b in zip(got, ref):
        assert rel(a, b) <= 2e-3          # atomics in the narrow / depthwise wgrad make the last bits order-dependent


def test_kernel_timing_facility_reports_library_kernels():
    import ctypes
    from quan_ultralytics_b200 import _lib
    lib = _lib.load()
    x = torch.randn(2, 64, 16, 16, 4, device=DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
    w = [torch.randn(64, 64, 3, 3, device=DEV) * 0.05 for _ in range(4)]
    lib.quan_kernel_timing_enable(1)
    for _ in range(3):
        ops.qconv2d_fwd(x, w, None, (1, 1), (1, 1), (1, 1), 1, ops.M_A, ops.ALGO_AUTO, L)
    torch.cuda.synchronize()
    lib.quan_kernel_timing_enable(0)
    n = lib.quan_kernel_timing_report(None, 0)
    buf = ctypes.create_string_buffer(n + 8)
    lib.quan_kernel_timing_report(buf, n + 8)
    rows = {ln.split()[0]: (int(ln.split()[1]), float(ln.split()[2])) for ln in buf.value.decode().splitlines()}
    assert rows["qconv_igemm_fwd"][0] == 3 and rows["qconv_igemm_fwd"][1] > 0.0
    assert rows["pack_weights_kernel"][0] == 3


@pytest.mark.parametrize("cfg", [("sep_bf16", torch.bfloat16, 64, 128, 1, 3), ("sep_f32", torch.float32, 64, 64, 1, 3),
                                 ("dense_bf16", torch.bfloat16, 16, 32, 1, 1), ("dw_bf16", torch.bfloat16, 32, 32, 32, 3),
                                 ("small_bf16", torch.bfloat16, 4, 2, 1, 3)], ids=lambda c: c[0])
def test_inference_conv_block_epilogue_fusion(cfg):
    """eval-mode `Conv` under no_grad (IQBN with running statistics + SiLU folded into the tensor-core epilogue, one C call)
    against the separate eval-mode nodes (conv.py:546-552 semantics, golden-checked elsewhere)."""
    import quan_ultralytics_b200 as Q
    name, dtype, ci, co, g, k = cfg
    torch.manual_seed(17)
    blk = (Q.DWConv(ci * 4, co * 4, k, 1) if g > 1 else Q.Conv(ci * 4, co * 4, k, 1)).to(DEV)
    with torch.no_grad():
        blk.bn.running_mean.normal_(0, 0.3)
        blk.bn.running_var.uniform_(0.5, 2.0)
        blk.bn.gamma.uniform_(0.5, 1.5)
        blk.bn.beta.normal_(0, 0.5)
    blk.eval()
    x = torch.randn(3, ci, 21, 19, 4, device=DEV).to(dtype).contiguous(memory_format=torch.channels_last_3d)
    with torch.no_grad():
        blk.fuse_block = True
        y_fused = blk(x)
        blk.fuse_block = False
        y_sep = blk(x)
    # the fused epilogue normalises the fp32 accumulator; the separate path rounds the conv output to bf16 first
    tol = 1.5e-2 if dtype == torch.bfloat16 else 1e-5
    assert y_fused.shape == y_sep.shape and rel(y_fused, y_sep) <= tol
    # against the oracle in fp64 on the same inputs
    xn = x.double().cpu().numpy()
    n64 = lambda t: t.detach().double().cpu().numpy()
    wn = [n64(getattr(blk.conv, f"weight_{c}")) for c in "rijk"]
    if dtype == torch.bfloat16:
        wn = [n64(getattr(blk.conv, f"weight_{c}").detach().to(torch.bfloat16)) for c in "rijk"]
    s_ = O.qconv2d_fwd(xn, wn, None, 1, k // 2, 1, g, O.M_A)
    y_ref = O.iqbn_eval_fwd(s_, n64(blk.bn.gamma), n64(blk.bn.beta), n64(blk.bn.running_mean), n64(blk.bn.running_var), act=True)
    err = float(np.max(np.abs(y_fused.double().cpu().numpy() - y_ref)) / np.max(np.abs(y_ref)))
    assert err <= (1e-2 if dtype == torch.bfloat16 else 1e-3)
