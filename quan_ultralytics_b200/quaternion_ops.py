"""Drop-in for the reference's compiled extension module `quaternion_ops`
(ultralytics/nn/cuda/quaternion_ops_py.cpp:132-165): same three entry points, same argument order and meaning,
same contiguous-BCHWQ tensor contract, RuntimeError on bad input — implemented on libquan_sm100.so.

Put this file's directory (or a `quaternion_ops.py` re-export) ahead of
`ultralytics/nn/cuda` on sys.path and the unmodified `conv.py:47-60` imports it and flips `CUDA_EXT = True`.

Mixing matrix: the reference extension computes M_B (quaternion_ops.cu:152-155) while the PyTorch path of the same
module computes M_A (conv.py:493-496) — SURVEY §0.1.  `set_mixing("B")` (default) is faithful to the extension;
`set_mixing("A")` makes the YOLO models reproduce their own PyTorch path, which is what BASELINE.json asks parity with.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from ._lib import ACT_NONE, ALGO_AUTO, ALGO_TCGEN05, LAYOUT_BCHWQ, LAYOUT_BHWQC

_mix = {"name": "B"}


def set_mixing(name: str) -> None:
    if name not in ops.MIX:
        raise ValueError("mixing must be 'A' or 'B'")
    _mix["name"] = name


def get_mixing() -> str:
    return _mix["name"]


# The extension's contract is contiguous BCHWQ in and out.  Large layers are still worth two layout conversions: the
# tensor-core engine (BHWQC, tf32 MMA for fp32 tensors: within BASELINE.json's 1e-3) does the bench layer in 0.23 ms where
# the CUDA-core engine that works on BCHWQ directly needs 5.3 ms (the reference's own kernels: 14.7 ms).  Small tensors
# and shapes the tensor-core engine does not take stay on the exact-fp32 CUDA-core path.
_fast = {"on": True, "min_numel": 1 << 18}


def set_fast_layout(on: bool, min_numel: int = 1 << 18) -> None:
    _fast["on"], _fast["min_numel"] = bool(on), int(min_numel)


def _tensor_core_ok(x: torch.Tensor, w: torch.Tensor, stride, padding, dilation, groups, passes) -> bool:
    if not _fast["on"] or x.numel() < _fast["min_numel"] or x.dtype not in (torch.float32, torch.bfloat16):
        return False
    return all(ops.qconv2d_pick_algo(x.shape, w.shape, tuple(stride), tuple(padding), tuple(dilation), int(groups), x.dtype,
                                     LAYOUT_BHWQC, ps) == ALGO_TCGEN05 for ps in passes)


def _from_half(t: torch.Tensor) -> torch.Tensor:
    """fp16 at the boundary: the reference's `QConvFunction` casts its inputs to float16 under autocast
    (quaternion_autograd_cuda.py:19, `custom_fwd(cast_inputs=torch.float16)`) and its kernels dispatch on half
    (quaternion_ops.cu:780).  fp16 values are widened to fp32 storage and run on the fp32 path: the tensor core's tf32 operand
    keeps 10 mantissa bits — exactly fp16's — so no input bit is lost (bf16 would drop three), accumulation is fp32 as in the
    reference kernels, and results are rounded back to fp16 on the way out."""
    return t.float() if t is not None and t.dtype == torch.float16 else t


def _check_cuda(name: str, t: torch.Tensor) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")          # quaternion_ops_py.cpp:69-81


def qconv_forward(input: torch.Tensor, weight_r, weight_i, weight_j, weight_k, bias_r: Optional[torch.Tensor],
                  bias_i, bias_j, bias_k, stride: Sequence[int], padding: Sequence[int], dilation: Sequence[int],
                  groups: int) -> torch.Tensor:
    """-> Tensor[B, C_out, H_out, W_out, 4] (quaternion_ops_py.cpp:48-86)."""
    _check_cuda("Input", input)
    for n, w in zip("rijk", (weight_r, weight_i, weight_j, weight_k)):
        _check_cuda(f"weight_{n}", w)
    if bias_r is None and not (bias_i is None and bias_j is None and bias_k is None):
        raise RuntimeError("If bias_r is None, bias_i, bias_j, and bias_k must also be None.")
    if bias_r is not None:
        _check_cuda("bias_r", bias_r)
    if input.dtype == torch.float16:
        y = qconv_forward(input.float(), _from_half(weight_r), _from_half(weight_i), _from_half(weight_j), _from_half(weight_k),
                          _from_half(bias_r), None, None, None, stride, padding, dilation, groups)
        return y.half()
    x = input.contiguous()
    if _tensor_core_ok(x, weight_r, stride, padding, dilation, groups, (0,)):
        y = ops.qconv2d_fwd(ops.convert_layout(x, LAYOUT_BHWQC), (weight_r, weight_i, weight_j, weight_k), bias_r, tuple(stride),
                            tuple(padding), tuple(dilation), int(groups), ops.MIX[_mix["name"]], ALGO_AUTO, LAYOUT_BHWQC)
        return ops.convert_layout(y, LAYOUT_BCHWQ)
    y = ops.qconv2d_fwd(x, (weight_r, weight_i, weight_j, weight_k), bias_r, tuple(stride),
                        tuple(padding), tuple(dilation), int(groups), ops.MIX[_mix["name"]], ALGO_AUTO, LAYOUT_BCHWQ)
    return y


def qconv_backward(grad_output: torch.Tensor, input: torch.Tensor, weight_r, weight_i, weight_j, weight_k,
                   bias_defined: bool, stride, padding, dilation, groups: int) -> List[Optional[torch.Tensor]]:
    """-> [dX, dW_r, dW_i, dW_j, dW_k, db_r or None] (quaternion_ops_py.cpp:89-111).
    db_r is the autograd-correct sum_p M[p,0]·dY_p, not the reference kernel's raw sum dY_r (SURVEY §8(c) defect 3)."""
    _check_cuda("grad_output", grad_output)
    _check_cuda("Input", input)
    if input.dtype == torch.float16:
        ws16 = (weight_r, weight_i, weight_j, weight_k)
        out = qconv_backward(grad_output.float(), input.float(), *[_from_half(w) for w in ws16], bias_defined, stride, padding,
                             dilation, groups)
        return [out[0].half()] + [g.to(w.dtype) for g, w in zip(out[1:5], ws16)] + [None if out[5] is None else out[5].to(weight_r.dtype)]
    x = input.contiguous()
    dy = grad_output.contiguous()
    if dy.dtype != x.dtype:
        dy = dy.to(x.dtype)
    fast = _tensor_core_ok(x, weight_r, stride, padding, dilation, groups, (1, 2))
    if fast:
        x, dy = ops.convert_layout(x, LAYOUT_BHWQC), ops.convert_layout(dy, LAYOUT_BHWQC)
    dx, dws, db = ops.qconv2d_bwd(dy, x, (weight_r, weight_i, weight_j, weight_k), tuple(stride), tuple(padding),
                                  tuple(dilation), int(groups), ops.MIX[_mix["name"]], True, True, bool(bias_defined))
    if fast:
        dx = ops.convert_layout(dx, LAYOUT_BCHWQ)
    dws = [g.to(w.dtype) for g, w in zip(dws, (weight_r, weight_i, weight_j, weight_k))]
    if db is not None:
        db = db.to(weight_r.dtype)
    return [dx, dws[0], dws[1], dws[2], dws[3], db]


def iqbn_forward(input: torch.Tensor, gamma, beta, running_mean, running_var, eps: float) -> torch.Tensor:
    """Eval-mode IQBN (quaternion_ops_py.cpp:113-128 -> quaternion_ops.cu:8-39)."""
    for n, t in (("Input", input), ("Gamma", gamma), ("Beta", beta), ("Running mean", running_mean),
                 ("Running variance", running_var)):
        _check_cuda(n, t)
    if input.dtype == torch.float16:
        return iqbn_forward(input.float(), gamma, beta, running_mean, running_var, eps).half()
    x, layout = ops.as_layout(input)
    f = ops._f32c
    return ops.iqbn_eval_fwd(x, layout, f(gamma), f(beta), f(running_mean), f(running_var), float(eps), ACT_NONE)
