"""Host-side mirror of the reference nn.Module API: names, shapes, state-dict keys, padding rules, init statistics,
and the no-CPU-fallback rule.  (Compute parity lives in the `-m gpu` tests.)"""
import math

import pytest
import torch

import quan_ultralytics_b200 as Q
from quan_ultralytics_b200 import ops


def test_qconv2d_parameters_match_reference_layout():
    m = Q.QConv2D(64, 128, 3, stride=2, padding=1, groups=2, bias=True)
    assert m.weight_r.shape == (32, 8, 3, 3) == m.weight_i.shape == m.weight_j.shape == m.weight_k.shape
    assert m.bias_r.shape == (32,) and m.bias_i is None and m.bias_j is None and m.bias_k is None
    keys = set(m.state_dict().keys())
    assert keys == {"weight_r", "weight_i", "weight_j", "weight_k", "bias_r"}       # conv.py:139-151
    assert m.stride == (2, 2) and m.padding == (1, 1) and m.dilation == (1, 1)
    first = Q.QConv2D(3, 16, 3)
    assert first.is_first_layer and first.in_channels_per_comp == 1 and first.weight_r.shape == (4, 1, 3, 3)


def test_padding_same_rules():
    assert Q.QConv2D(16, 16, 5, padding="same").padding == (2, 2)                     # conv.py:94-95
    assert Q.QConv2D(16, 16, 3, stride=2, padding="same", dilation=2).padding == (2, 2)   # autopad with dilation
    with pytest.raises(ValueError):
        Q.QConv2D(16, 16, 3, padding="bogus")
    with pytest.raises(AssertionError):
        Q.QConv2D(6, 16, 3)          # in_channels must be a multiple of 4 (or 3 for RGB)


def test_init_matches_kaiming_uniform_a_sqrt5():
    torch.manual_seed(0)
    m = Q.QConv2D(256, 256, 3, bias=True)
    fan_in = 64 * 9
    bound = math.sqrt(6.0 / ((1 + 5) * fan_in))                   # kaiming_uniform_(a=sqrt(5)) bound
    for w in (m.weight_r, m.weight_i, m.weight_j, m.weight_k):
        assert w.abs().max() <= bound + 1e-7
        assert abs(w.std().item() - bound / math.sqrt(3)) < 0.02 * bound
    assert m.bias_r.abs().max() <= 1 / math.sqrt(fan_in) + 1e-7


def test_iqbn_buffers_and_shapes():
    bn = Q.IQBN(64)
    assert bn.num_features == 16
    sd = bn.state_dict()
    assert set(sd) == {"gamma", "beta", "running_mean", "running_var", "num_batches_tracked"}   # conv.py:512-518
    assert sd["gamma"].shape == (16, 4) and sd["running_var"].shape == (16, 4)
    with pytest.raises(AssertionError):
        Q.IQBN(6)


def test_conv_and_dwconv_wrappers():
    c = Q.Conv(64, 128, 3, 2)
    assert isinstance(c.conv, Q.QConv2D) and c.conv.bias_r is None and c.conv.padding == (1, 1)
    assert isinstance(c.bn, Q.IQBN) and isinstance(c.act, torch.nn.SiLU)
    dw = Q.DWConv(64, 64, 3)
    assert dw.conv.groups == 16 and dw.conv.weight_r.shape == (16, 1, 3, 3)       # conv.py:918-923
    assert isinstance(Q.Conv(16, 16, act=False).act, torch.nn.Identity)


def test_mix_selection():
    assert Q.QConv2D(16, 16, 1).mix == "A" and Q.QConv2D_B(16, 16, 1).mix == "B"
    assert ops.M_A[:4] == (1., -1., -1., -1.) and ops.M_B[4:8] == (1., -1., -1., 1.)
    assert ops._mix_t(ops.M_A)[1] == ops.M_A[4]


def test_no_cpu_fallback():
    m = Q.QConv2D(16, 16, 3, padding=1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 4, 8, 8, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Q.QUpsample()(torch.randn(1, 4, 8, 8, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Q.poincare_map(torch.rand(1, 3, 8, 8))


def test_layout_detection():
    x = torch.randn(2, 8, 4, 4, 4)
    assert ops.layout_of(x) == ops.LAYOUT_BCHWQ
    assert ops.layout_of(x.contiguous(memory_format=torch.channels_last_3d)) == ops.LAYOUT_BHWQC
    assert ops.layout_of(x[:, ::2]) is None
    assert ops.layout_of(torch.randn(2, 1, 4, 4, 4)) == ops.LAYOUT_BCHWQ
    with pytest.raises(RuntimeError):
        ops.layout_of(torch.randn(2, 8, 4, 4))


def test_unsupported_options_fail_loudly():
    with pytest.raises(NotImplementedError):
        Q.QUpsample(2, "bilinear")
    with pytest.raises(NotImplementedError):
        Q.QConv2D(16, 16, 3, padding_mode="reflect")


def test_qer_reads_the_tensor_core_layout_without_a_copy():
    """QER (head.py:26-47): in BHWQC the activation is already a channels-last real tensor; only the weight is re-ordered.
    Pure torch ops, so it is checked on CPU against the reference's formulation (permute + contiguous + view + conv)."""
    import quan_ultralytics_b200 as Q
    torch.manual_seed(0)
    B, C, H, W, O = 2, 6, 5, 7, 9
    m = Q.QER(4 * C, O, 1).double()
    assert set(m.state_dict()) == {"bias", "output_proj.weight", "output_proj.bias"}
    x = torch.randn(B, C, H, W, 4, dtype=torch.float64)
    ref_in = x.clone().requires_grad_(True)
    ref = m.output_proj(ref_in.permute(0, 1, 4, 2, 3).contiguous().view(B, C * 4, H, W))
    dy = torch.randn_like(ref)
    ref.backward(dy)
    gw, gb = m.output_proj.weight.grad.clone(), m.output_proj.bias.grad.clone()
    for fmt in (torch.channels_last_3d, torch.contiguous_format):
        m.zero_grad(set_to_none=True)
        xi = x.clone().contiguous(memory_format=fmt).requires_grad_(True)
        if fmt == torch.channels_last_3d:          # the fast branch really is a view of the activation
            v = xi.detach().permute(0, 4, 1, 2, 3).reshape(B, 4 * C, H, W)
            assert v.data_ptr() == xi.data_ptr() and v.is_contiguous(memory_format=torch.channels_last)
        y = m(xi)
        assert y.shape == (B, O, H, W)
        torch.testing.assert_close(y, ref, rtol=1e-12, atol=1e-12)
        y.backward(dy)
        torch.testing.assert_close(xi.grad, ref_in.grad, rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(m.output_proj.weight.grad, gw, rtol=1e-12, atol=1e-12)
        torch.testing.assert_close(m.output_proj.bias.grad, gb, rtol=1e-12, atol=1e-12)
