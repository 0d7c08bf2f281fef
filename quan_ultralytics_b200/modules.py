"""nn.Module mirror of the reference's quaternion layer API, computing through libquan_sm100.so.

Same class names, constructor signatures, parameter/buffer names and shapes as
  ultralytics/nn/modules/conv.py      — QConv2D :70-499, IQBN :501-571, Conv :788-813, DWConv :918-923, QUpsample :1218-1246
  classification/quaternion/qconv.py  — QConv2D :399-612 (mixing matrix M_B), IQBN :337-396
so state dicts are interchangeable (weight_r/i/j/k, bias_r, gamma, beta, running_mean, running_var,
num_batches_tracked — SURVEY §5 "Checkpoint / resume").  Differences, all deliberate and documented in DESIGN.md:
  * IQBN uses batch statistics whenever `self.training` (the reference's `not CUDA_EXT` guard, conv.py:537, is a
    defect — SURVEY §0.2) and updates the running buffers in place instead of re-assigning them.
  * `mix` selects the reference's mixing matrix: "A" (ultralytics conv.py:493-496) or "B" (classification/ext).
  * tensors stay in the internal BHWQC (channels_last_3d) layout between layers; logical shapes are unchanged.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import functional as QF
from . import ops
from ._lib import ACT_NONE, ACT_SILU, ALGO_AUTO

__all__ = ["QConv2D", "QConv2D_B", "IQBN", "QUpsample", "Conv", "DWConv", "autopad"]


def autopad(k, p=None, d=1):
    """Pad to 'same' shape outputs (conv.py:62-68)."""
    if d > 1:
        k = d * (k - 1) + 1 if isinstance(k, int) else [d * (x - 1) + 1 for x in k]
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


class QConv2D(nn.Module):
    """Separable Hamilton-product convolution, y = M · (W_sep * x) (conv.py:70-499)."""

    default_mix = "A"

    def __init__(self, in_channels: int, out_channels: int, kernel_size: Union[int, Tuple[int, int]],
                 stride: Union[int, Tuple[int, int]] = 1, padding: Union[str, int, Tuple[int, int]] = 0,
                 dilation: Union[int, Tuple[int, int]] = 1, groups: int = 1, bias: bool = True,
                 padding_mode: str = "zeros", dtype=None, mapping_type: str = "poincare",
                 mix: Optional[str] = None):
        super().__init__()
        if isinstance(kernel_size, int):
            kernel_size = (kernel_size, kernel_size)
        if isinstance(stride, int):
            stride = (stride, stride)
        if isinstance(dilation, int):
            dilation = (dilation, dilation)
        if padding == "same":  # conv.py:91-98
            if stride == (1, 1):
                padding = ((kernel_size[0] - 1) // 2, (kernel_size[1] - 1) // 2)
            else:
                padding = tuple(autopad(k, None, d) for k, d in zip(kernel_size, dilation))
        elif isinstance(padding, int):
            padding = (padding, padding)
        elif isinstance(padding, (list, tuple)) and len(padding) == 2:
            padding = tuple(padding)
        else:
            raise ValueError(f"Invalid padding: {padding}")
        if padding_mode != "zeros":
            raise NotImplementedError("only padding_mode='zeros' is implemented (the reference ignores the argument)")

        self.in_channels_total = in_channels
        self.out_channels_total = out_channels
        self.kernel_size = tuple(kernel_size)
        self.stride = tuple(stride)
        self.padding = padding
        self.dilation = tuple(dilation)
        self.groups = groups
        self.mapping_type = mapping_type
        self.mix = mix or self.default_mix
        self.algo = ALGO_AUTO

        self.is_first_layer = in_channels == 3
        if self.is_first_layer:
            self.in_channels_per_comp = 1
        else:
            assert in_channels % 4 == 0, "in_channels must be multiple of 4 for non-first layers"
            self.in_channels_per_comp = in_channels // 4
        assert out_channels % 4 == 0, "out_channels must be multiple of 4"
        self.out_channels_per_comp = out_channels // 4
        assert self.in_channels_per_comp % groups == 0, "Input channels per component must be divisible by groups"
        self.in_channels_per_comp_grp = self.in_channels_per_comp // groups
        self.bias_flag_overall = bias

        shape = (self.out_channels_per_comp, self.in_channels_per_comp_grp, *self.kernel_size)
        self.weight_r = nn.Parameter(torch.zeros(shape))
        self.weight_i = nn.Parameter(torch.zeros(shape))
        self.weight_j = nn.Parameter(torch.zeros(shape))
        self.weight_k = nn.Parameter(torch.zeros(shape))
        if bias:
            self.bias_r = nn.Parameter(torch.zeros(self.out_channels_per_comp))
        else:
            self.register_parameter("bias_r", None)
        self.register_parameter("bias_i", None)
        self.register_parameter("bias_j", None)
        self.register_parameter("bias_k", None)
        self._initialize_weights()

    def _initialize_weights(self):
        """kaiming_uniform(a=sqrt(5)·scale) per component, bias U(±1/sqrt(fan_in)) (conv.py:232-256)."""
        fan_in = self.in_channels_per_comp_grp * int(np.prod(self.kernel_size))
        scales = {"mean_brightness": [1.0, 0.75, 0.75, 0.75]}.get(
            self.mapping_type,
            [1.0] * 4 if self.mapping_type in ("luminance", "raw_normalized", "hamilton", "poincare") else [0.5] * 4)
        for i, w in enumerate((self.weight_r, self.weight_i, self.weight_j, self.weight_k)):
            nn.init.kaiming_uniform_(w, a=math.sqrt(5.0) * scales[i])
        if self.bias_flag_overall and self.bias_r is not None:
            bound = (1 / math.sqrt(fan_in)) * scales[0] if fan_in > 0 else 0
            nn.init.uniform_(self.bias_r, -bound, bound)

    def _rgb_to_quaternion(self, rgb: torch.Tensor) -> torch.Tensor:
        if self.mapping_type != "poincare":
            raise NotImplementedError(
                f"mapping_type={self.mapping_type!r}: only the Poincare map (conv.py:388-397) is part of the "
                "B200 hot path (SURVEY §8 a3)")
        return QF.poincare_map(rgb)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.is_first_layer:
            assert x.dim() == 4 and x.size(1) == 3, f"Expected [B,3,H,W] RGB input, got {tuple(x.shape)}"
            x = self._rgb_to_quaternion(x)
        elif x.dim() == 4:  # [B, 4C, H, W] -> BCHWQ view (conv.py:427-433)
            B, C, H, W = x.shape
            assert C == self.in_channels_total, f"Expected {self.in_channels_total} channels, got {C}"
            x = x.view(B, C // 4, 4, H, W).permute(0, 1, 3, 4, 2)
        elif x.dim() == 5:
            assert x.size(1) == self.in_channels_per_comp, \
                f"Input C_per_q mismatch {x.size(1)} vs {self.in_channels_per_comp}"
            assert x.size(4) == 4, "Input quaternion dim must be 4"
        else:
            raise ValueError(f"Unsupported input shape: {x.shape}")
        return QF.qconv2d(x, self.weight_r, self.weight_i, self.weight_j, self.weight_k, self.bias_r, self.stride,
                          self.padding, self.dilation, self.groups, ops.MIX[self.mix], self.algo)

    def extra_repr(self) -> str:
        return (f"{self.in_channels_total}, {self.out_channels_total}, kernel_size={self.kernel_size}, "
                f"stride={self.stride}, padding={self.padding}, groups={self.groups}, "
                f"bias={self.bias_r is not None}, mix={self.mix}")


class QConv2D_B(QConv2D):
    """classification/quaternion/qconv.py:399-612 flavour (mixing matrix M_B, also what the reference's CUDA
    extension computes, quaternion_ops.cu:152-155)."""

    default_mix = "B"


class IQBN(nn.Module):
    """Independent quaternion batch-norm over (B,H,W) per (channel, component) (conv.py:501-571)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        assert num_features % 4 == 0, "num_features must be a multiple of 4 for IQBN"
        self.num_features = num_features // 4
        self.eps = eps
        self.momentum = momentum
        self.gamma = nn.Parameter(torch.ones(self.num_features, 4))
        self.beta = nn.Parameter(torch.zeros(self.num_features, 4))
        self.register_buffer("running_mean", torch.zeros(self.num_features, 4))
        self.register_buffer("running_var", torch.ones(self.num_features, 4))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))
        self.process_group = None   # set by distributed.convert_sync_iqbn
        self.sync = False
        self._nbt_pooled = False    # set by pool_batch_counters: the model's forward pre-hook counts for every IQBN at once

    def forward(self, x: torch.Tensor, act: int = ACT_NONE) -> torch.Tensor:
        assert x.dim() == 5 and x.size(4) == 4, "Input must be [B, C, H, W, 4]"
        assert x.size(1) == self.num_features, f"expected {self.num_features} quaternion channels, got {x.size(1)}"
        if self.training:
            if not self._nbt_pooled:
                with torch.no_grad():
                    self.num_batches_tracked += 1
            group = None
            if self.sync and torch.distributed.is_available() and torch.distributed.is_initialized():
                group = self.process_group or torch.distributed.group.WORLD
            return QF.iqbn(x, self.gamma, self.beta, self.running_mean, self.running_var, True, self.eps,
                           self.momentum, act, group)
        return QF.iqbn(x, self.gamma, self.beta, self.running_mean, self.running_var, False, self.eps,
                       self.momentum, act)

    def extra_repr(self) -> str:
        return f"{self.num_features * 4}, eps={self.eps}, momentum={self.momentum}, sync={self.sync}"


def pool_batch_counters(model: nn.Module) -> Optional[torch.Tensor]:
    """`num_batches_tracked += 1` (conv.py:563) is one tiny kernel per IQBN per training forward — 84 launches of the QUAN-YOLO11n
    step.  This re-homes every IQBN's counter as a view into ONE int64 tensor (same buffer names, shapes and state-dict entries) and
    counts with a single add in a forward pre-hook of `model` (training mode only).  Call it after the model sits on its device;
    a later `.to(other_device)` gives every buffer its own storage again — call it again then.  Counts every pooled IQBN on every
    training forward of `model` (the reference counts the ones that ran — all of them, in the QUAN graphs)."""
    bns = [m for m in model.modules() if isinstance(m, IQBN)]
    if not bns:
        return None
    dev = bns[0].num_batches_tracked.device
    pool = torch.stack([m.num_batches_tracked.detach().to(dev, torch.long).reshape(()) for m in bns])
    for i, m in enumerate(bns):
        m._buffers["num_batches_tracked"] = pool[i]
        m._nbt_pooled = True
    old = getattr(model, "_quan_nbt_hook", None)
    if old is not None:
        old.remove()
    model._quan_nbt_pool = pool

    def count(mod, args):
        if mod.training:
            mod._quan_nbt_pool.add_(1)

    model._quan_nbt_hook = model.register_forward_pre_hook(count)
    return pool


class QUpsample(nn.Module):
    """Nearest-neighbour upsampling of every quaternion component (conv.py:1218-1246)."""

    def __init__(self, scale_factor=2, mode="nearest"):
        super().__init__()
        if mode != "nearest":
            raise NotImplementedError("QUpsample: only mode='nearest' is used by the QUAN models (yolo11-obb-quan.yaml:35,39)")
        if int(scale_factor) != scale_factor:
            raise NotImplementedError("QUpsample: integer scale factors only")
        self.scale_factor = int(scale_factor)
        self.mode = mode

    def forward(self, x):
        assert x.dim() == 5 and x.size(4) == 4, "Expected quaternion format [B, C, H, W, 4]"
        return QF.qupsample_nearest(x, self.scale_factor)


class QuaternionMaxPool(nn.Module):
    """Max pooling of every quaternion component (ultralytics/nn/modules/block.py:85-109 ==
    classification/models/blocks/quaternion_blocks.py:236-260); constructor as the reference's."""

    def __init__(self, kernel_size=2, stride=2, padding=0):
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.dim() == 5 and x.size(4) == 4, "Expected quaternion format with 4 components"
        return QF.qmaxpool(x, self.kernel_size, self.stride, self.padding)

    def extra_repr(self) -> str:
        return f"kernel_size={self.kernel_size}, stride={self.stride}, padding={self.padding}"


class QER(nn.Module):
    """Quaternion -> real extraction of the detection heads (ultralytics/nn/modules/head.py:26-47): a real `nn.Conv2d` over
    the 4C flattened quaternion channels (channel index c*4 + q).  The reference first materialises
    `x.permute(0, 1, 4, 2, 3).contiguous()`; in the tensor-core layout (BHWQC) the activation already IS a channels-last
    real tensor with channel index q*C + c, so the copy disappears and the projection is quan_qer_fwd / quan_qer_bwd
    (csrc/qer.cu: pixel-row GEMM with the weight re-ordered while it is staged, bias and bias gradient fused).  Same
    constructor, parameters and state-dict keys (`output_proj.weight`, `output_proj.bias`, `bias`) as the reference; output is
    the same logical [B, out, H, W] tensor in channels-last memory.  Shapes outside the kernel's range (4C > 256, N > 64, a
    kernel size other than 1) and CPU tensors take the view + re-ordered weight + library conv path.  SURVEY §8(f) rank 2."""

    def __init__(self, in_channels, out_channels=None, kernel_size=None):
        super().__init__()
        self.output_proj = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size)
        self.bias = self.output_proj.bias

    def fused_ok(self, x: torch.Tensor) -> bool:
        """True when quan_qer_* serves this call: a CUDA [B,C,H,W,4] activation, the plain 1x1 projection, supported widths."""
        conv = self.output_proj
        return (x.dim() == 5 and ops.on_device(x) and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.padding == (0, 0)
                and conv.groups == 1 and conv.padding_mode == "zeros" and ops.qer_supported(x.size(1), conv.out_channels, x.dtype))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, C, H, W, Q = x.shape
        conv = self.output_proj
        if self.fused_ok(x):
            return QF.qer(x, conv.weight, conv.bias)             # the library's kernel: reads BHWQC rows, fp32 master weight
        if C > 1 and x.is_contiguous(memory_format=torch.channels_last_3d):
            xr = x.permute(0, 4, 1, 2, 3).reshape(B, Q * C, H, W)            # a view: memory is [B][H][W][q*C + c]
            w = conv.weight
            w = w.view(w.size(0), C, Q, w.size(2), w.size(3)).transpose(1, 2).reshape(w.size(0), Q * C, w.size(2), w.size(3))
            return torch.nn.functional.conv2d(xr, w, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
        return conv(x.permute(0, 1, 4, 2, 3).contiguous().view(B, C * Q, H, W))  # BCHWQ input: the reference's own path


class Conv(nn.Module):
    """QConv2D -> IQBN -> SiLU (conv.py:788-813); IQBN-apply and SiLU run as one fused kernel."""

    default_act = nn.SiLU()
    conv_cls = QConv2D
    fuse_block = True      # training: run conv + IQBN + act as one autograd node (functional.conv_iqbn_act)

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = self.conv_cls(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = IQBN(c2)
        self.act = self.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()

    def forward(self, x):
        fused_act = ACT_SILU if isinstance(self.act, nn.SiLU) else ACT_NONE if isinstance(self.act, nn.Identity) else None
        c, bn = self.conv, self.bn
        if (self.fuse_block and fused_act is not None and bn.training and not bn.sync and c.bias_r is None
                and not c.is_first_layer and x.dim() == 5 and ops.on_device(x)):
            # conv -> IQBN -> act as one autograd node: the IQBN backward hands G = M^T dY straight to the conv backward
            if not bn._nbt_pooled:
                with torch.no_grad():
                    bn.num_batches_tracked += 1
            return QF.conv_iqbn_act(x, c.weight_r, c.weight_i, c.weight_j, c.weight_k, bn.gamma, bn.beta, bn.running_mean,
                                    bn.running_var, c.stride, c.padding, c.dilation, c.groups, ops.MIX[c.mix], c.algo,
                                    bn.eps, bn.momentum, fused_act)
        if (self.fuse_block and fused_act is not None and not bn.training and c.bias_r is None and not c.is_first_layer
                and x.dim() == 5 and ops.on_device(x) and not torch.is_grad_enabled()):
            # inference (eval mode under no_grad): one C call; on the tensor-core engine IQBN(running stats) + act run in the
            # conv epilogue.  With autograd enabled the separate nodes below keep every gradient path.
            xl, layout = ops.as_layout(x, QF.internal_layout())
            return ops.conv_block_eval_fwd(xl, (c.weight_r, c.weight_i, c.weight_j, c.weight_k), ops._f32c(bn.gamma),
                                           ops._f32c(bn.beta), ops._f32c(bn.running_mean), ops._f32c(bn.running_var), c.stride,
                                           c.padding, c.dilation, c.groups, ops.MIX[c.mix], c.algo, bn.eps, fused_act, layout)
        y = self.conv(x)
        if isinstance(self.act, nn.SiLU):
            return self.bn(y, ACT_SILU)
        return self.act(self.bn(y))

    def forward_fuse(self, x):
        return self.act(self.conv(x))


class DWConv(Conv):
    """Depth-wise quaternion convolution (conv.py:918-923)."""

    def __init__(self, c1, c2, k=1, s=1, d=1, act=True):
        super().__init__(c1, c2, k, s, g=math.gcd(c1 // 4, c2 // 4), d=d, act=act)
