"""CPU ORACLE (torch port) — TEST / BASELINE INFRASTRUCTURE ONLY.

A call-by-call restatement of the reference's PyTorch path with the same torch operators it uses (4x F.conv2d on
strided component slices + elementwise mix + torch.stack; Tensor.mean/var batch-norm; F.interpolate), so that it
(a) cross-checks oracle/quan_oracle.py, and (b) is the thing bench.py times as the reference's CPU path
(`cpu_baseline.kind = "port"`, `--impl reference`).  /root/reference cannot travel to the GPU box; this port can.
The product never imports this file.

Pinned against the real reference by tests/golden/*.npz (tests/test_oracle_golden.py).
Citations are reference-repo paths.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

M_A = torch.tensor([[1., -1, -1, -1], [-1, 1, 1, -1], [-1, -1, 1, 1], [-1, 1, -1, 1]])   # conv.py:493-496
M_B = torch.tensor([[1., 1, 1, 1], [1, -1, -1, 1], [1, 1, -1, -1], [1, -1, 1, -1]])       # qconv.py:606-609


def poincare(rgb: torch.Tensor) -> torch.Tensor:
    """conv.py:388-397."""
    norm_sq = torch.sum(rgb ** 2, dim=1, keepdim=True)
    den = 1 + norm_sq
    real = (1 - norm_sq.squeeze(1)) / den.squeeze(1)
    vec = 2 * rgb / den
    return torch.stack([real, vec[:, 0], vec[:, 1], vec[:, 2]], dim=-1).unsqueeze(1)


def qconv2d(x, w_r, w_i, w_j, w_k, bias_r, stride, padding, dilation, groups, mix: str):
    """conv.py:472-499 (mix='A') / classification/quaternion/qconv.py:592-612 (mix='B')."""
    p = dict(stride=stride, padding=padding, dilation=dilation, groups=groups)
    r = F.conv2d(x[..., 0], w_r, bias_r, **p)
    i = F.conv2d(x[..., 1], w_i, None, **p)
    j = F.conv2d(x[..., 2], w_j, None, **p)
    k = F.conv2d(x[..., 3], w_k, None, **p)
    if mix == "A":
        out = (r - i - j - k, -r + i + j - k, -r - i + j + k, -r + i - j + k)
    else:
        out = (r + i + j + k, r - i - j + k, r + i - j - k, r - i + j - k)
    return torch.stack(out, dim=-1)


class QConv2D(nn.Module):
    def __init__(self, cin_q, cout_q, k, stride=1, padding=0, dilation=1, groups=1, bias=True, mix="A"):
        super().__init__()
        pair = lambda v: (v, v) if isinstance(v, int) else tuple(v)
        self.k, self.stride, self.padding, self.dilation = pair(k), pair(stride), pair(padding), pair(dilation)
        self.groups, self.mix = groups, mix
        shape = (cout_q, cin_q // groups, *self.k)
        for n in "rijk":
            w = nn.Parameter(torch.zeros(shape))
            nn.init.kaiming_uniform_(w, a=math.sqrt(5.0))               # conv.py:250
            setattr(self, f"weight_{n}", w)
        if bias:
            fan_in = (cin_q // groups) * self.k[0] * self.k[1]
            self.bias_r = nn.Parameter(torch.empty(cout_q).uniform_(-1 / math.sqrt(fan_in), 1 / math.sqrt(fan_in)))
        else:
            self.bias_r = None

    def forward(self, x):
        if x.dim() == 4 and x.size(1) == 3:
            x = poincare(x)
        return qconv2d(x.contiguous(), self.weight_r, self.weight_i, self.weight_j, self.weight_k, self.bias_r,
                       self.stride, self.padding, self.dilation, self.groups, self.mix)


class IQBN(nn.Module):
    """conv.py:501-571, batch-statistics branch when training (what the reference does once CUDA_EXT is true /
    classification/quaternion/qconv.py:378-396 always)."""

    def __init__(self, c_q, eps=1e-5, momentum=0.1):
        super().__init__()
        self.eps, self.momentum = eps, momentum
        self.gamma = nn.Parameter(torch.ones(c_q, 4))
        self.beta = nn.Parameter(torch.zeros(c_q, 4))
        self.register_buffer("running_mean", torch.zeros(c_q, 4))
        self.register_buffer("running_var", torch.ones(c_q, 4))

    def forward(self, x):
        C = x.size(1)
        v = lambda t: t.view(1, C, 1, 1, 4)
        if not self.training:
            return (x - v(self.running_mean)) / torch.sqrt(v(self.running_var) + self.eps) * v(self.gamma) + v(self.beta)
        mean = x.mean(dim=[0, 2, 3])
        var = x.var(dim=[0, 2, 3], unbiased=False) + 1e-8
        with torch.no_grad():
            self.running_mean = (1 - self.momentum) * self.running_mean + self.momentum * mean
            self.running_var = (1 - self.momentum) * self.running_var + self.momentum * var
        xn = (x - v(mean)) / torch.sqrt(v(var) + self.eps)
        return xn * v(self.gamma) + v(self.beta)


class Conv(nn.Module):
    """conv.py:788-809: SiLU(IQBN(QConv2D(x)))."""

    def __init__(self, cin_q, cout_q, k=1, s=1, g=1, mix="A", act=True):
        super().__init__()
        self.conv = QConv2D(cin_q, cout_q, k, s, k // 2, groups=g, bias=False, mix=mix)
        self.bn = IQBN(cout_q)
        self.act = nn.SiLU() if act else nn.Identity()

    def forward(self, x):
        return self.act(self.bn(self.conv(x)))


def qupsample(x, scale=2):
    """conv.py:1229-1246."""
    B, C, H, W, Q = x.shape
    flat = x.permute(0, 1, 4, 2, 3).reshape(B, C * Q, H, W)
    up = F.interpolate(flat, scale_factor=scale, mode="nearest")
    return up.view(B, C, Q, H * scale, W * scale).permute(0, 1, 3, 4, 2)
