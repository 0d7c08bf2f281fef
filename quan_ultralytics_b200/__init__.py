"""quan_ultralytics_b200 — B200 (sm_100a) implementation of QUAN's quaternion layer stack.

Scope: QConv2D separable Hamilton convolution, IQBN, QUpsample, the Poincare RGB->quaternion map (and, from the
§8(f) 'next' rows, QuaternionMaxPool),
forward and backward, behind the reference's nn.Module / `quaternion_ops` extension API.  Host code is
Python/PyTorch (memory, streams, autograd, torch.distributed); all arithmetic runs in hand-written CUDA behind the
C ABI in include/quan_sm100.h (libquan_sm100.so).  There is no CPU / PyTorch fallback.
"""
from . import _lib, ops  # noqa: F401
from ._lib import (ACT_NONE, ACT_SILU, ALGO_AUTO, ALGO_DEPTHWISE, ALGO_DIRECT, ALGO_TCGEN05, LAYOUT_BCHWQ,  # noqa: F401
                   LAYOUT_BHWQC)
from .functional import (conv_iqbn_act, iqbn, internal_layout, poincare_map, qconv2d, qmaxpool, qupsample_nearest,  # noqa: F401
                         set_epilogue_stats, set_internal_layout)
from .modules import IQBN, QER, Conv, DWConv, QConv2D, QConv2D_B, QuaternionMaxPool, QUpsample, autopad  # noqa: F401

__version__ = "0.1.0"
