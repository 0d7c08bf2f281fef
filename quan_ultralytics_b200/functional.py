"""torch.autograd bridge for the sm_100a quaternion ops.

Plays the role of the reference's `QConvFunction` (ultralytics/nn/modules/quaternion_autograd_cuda.py:18-75) and
extends it to the ops the reference left to PyTorch autograd (IQBN training fwd/bwd, QUpsample, the Poincare map).
All arithmetic happens in libquan_sm100.so; this file only wires saved tensors and gradients.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist

from . import ops
from ._lib import ACT_NONE, ACT_SILU, ALGO_AUTO

_state = {"layout": ops.LAYOUT_BHWQC, "epilogue_stats": True}


def set_epilogue_stats(on: bool) -> None:
    """Fused `Conv` node: take the IQBN batch statistics from the conv epilogue (default) or from a pass over y."""
    _state["epilogue_stats"] = bool(on)


def set_internal_layout(name: str) -> None:
    """'bhwqc' (channels_last_3d, tensor-core path; default) or 'bchwq' (the reference's contiguous layout)."""
    _state["layout"] = {"bhwqc": ops.LAYOUT_BHWQC, "bchwq": ops.LAYOUT_BCHWQ}[name.lower()]


def internal_layout() -> int:
    return _state["layout"]


class _Poincare(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb: torch.Tensor, out_dtype: torch.dtype):
        ctx.save_for_backward(rgb)
        return ops.poincare_fwd(rgb, out_dtype)

    @staticmethod
    def backward(ctx, grad_out):
        (rgb,) = ctx.saved_tensors
        grad = ops.poincare_bwd(rgb, grad_out) if ctx.needs_input_grad[0] else None
        if grad is not None and grad.dtype != rgb.dtype:
            grad = grad.to(rgb.dtype)
        return grad, None


def poincare_map(rgb: torch.Tensor, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """RGB [B,3,H,W] -> unit quaternions [B,1,H,W,4] (conv.py:388-397)."""
    if out_dtype is None:
        out_dtype = torch.bfloat16 if (rgb.dtype == torch.bfloat16 or
                                       (torch.is_autocast_enabled() and
                                        torch.get_autocast_dtype("cuda") == torch.bfloat16)) else torch.float32
    return _Poincare.apply(rgb, out_dtype)


class _QConv2d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_r, w_i, w_j, w_k, bias_r, stride, padding, dilation, groups, mix, algo):
        x, layout = ops.as_layout(x, _state["layout"])
        ctx.save_for_backward(x, w_r, w_i, w_j, w_k)
        ctx.conf = (tuple(stride), tuple(padding), tuple(dilation), int(groups), tuple(mix), algo, bias_r is not None)
        return ops.qconv2d_fwd(x, (w_r, w_i, w_j, w_k), bias_r, stride, padding, dilation, groups, mix, algo, layout)

    @staticmethod
    def backward(ctx, grad_out):
        x, w_r, w_i, w_j, w_k = ctx.saved_tensors
        stride, padding, dilation, groups, mix, algo, has_bias = ctx.conf
        need_dx = ctx.needs_input_grad[0]
        need_dw = any(ctx.needs_input_grad[1:5])
        need_db = has_bias and ctx.needs_input_grad[5]
        dx, dws, db = ops.qconv2d_bwd(grad_out, x, (w_r, w_i, w_j, w_k), stride, padding, dilation, groups, mix,
                                      need_dx, need_dw, need_db, algo)
        if dws is None:
            dws = [None] * 4
        else:
            dws = [g.to(w.dtype) if g.dtype != w.dtype else g for g, w in zip(dws, (w_r, w_i, w_j, w_k))]
        return (dx, dws[0], dws[1], dws[2], dws[3], db, None, None, None, None, None, None)


def qconv2d(x: torch.Tensor, w_r, w_i, w_j, w_k, bias_r=None, stride=(1, 1), padding=(0, 0), dilation=(1, 1),
            groups: int = 1, mix: Sequence[float] = ops.M_A, algo: int = ALGO_AUTO) -> torch.Tensor:
    """Separable Hamilton convolution on a BCHWQ tensor; same argument meaning as the reference's
    `qconv2d_function` (quaternion_autograd_cuda.py:72-75) plus the mixing matrix."""
    return _QConv2d.apply(x, w_r, w_i, w_j, w_k, bias_r, stride, padding, dilation, groups, mix, algo)


class _IQBNTrain(torch.autograd.Function):
    """Batch-statistics IQBN (+ optional fused SiLU).  Saves only x and the [12C] stats."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, act, group):
        x, layout = ops.as_layout(x)
        B, C_, H, W, _ = x.shape
        g32, b32 = ops._f32c(gamma), ops._f32c(beta)
        world = dist.get_world_size(group) if (group is not None and dist.is_initialized()) else 1
        if world > 1:  # synced IQBN: all-reduce {sum, sumsq} (new work)
            sums = ops.iqbn_partial_sums(x, layout)
            dist.all_reduce(sums, group=group)
            count = float(B * H * W * world)
            stats = ops.iqbn_finalize_stats(sums, count, C_, g32, b32, eps, momentum, running_mean, running_var)
        else:
            count = float(B * H * W)
            stats = ops.iqbn_train_stats(x, layout, g32, b32, eps, momentum, running_mean, running_var)
        y = ops.iqbn_apply_fwd(x, layout, stats, g32, b32, act)
        ctx.save_for_backward(x, stats, g32, b32)
        ctx.conf = (layout, act, count, group if world > 1 else None, gamma.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats, g32, b32 = ctx.saved_tensors
        layout, act, count, group, pdtype = ctx.conf
        dy, _ = ops.as_layout(dy, layout)
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        C_ = x.size(1)
        if group is not None:
            # parameter grads stay LOCAL sums (DDP averages them); dx needs the global sums
            sums = ops.iqbn_bwd_reduce(dy, x, layout, stats, g32, b32, act, 0.0)
            dbeta = sums[:4 * C_].to(torch.float32).view(C_, 4)
            dgamma = sums[4 * C_:8 * C_].to(torch.float32).view(C_, 4)
            dist.all_reduce(sums[:8 * C_], group=group)
            ops.iqbn_bwd_coef(sums, count, stats, g32)
            dx, _, _ = ops.iqbn_bwd_apply(dy, x, layout, stats, g32, b32, act, sums, count, want_param_grads=False)
        else:
            sums = ops.iqbn_bwd_reduce(dy, x, layout, stats, g32, b32, act, count)
            dx, dgamma, dbeta = ops.iqbn_bwd_apply(dy, x, layout, stats, g32, b32, act, sums, count)
        if pdtype != torch.float32:
            dgamma, dbeta = dgamma.to(pdtype), dbeta.to(pdtype)
        return dx, dgamma, dbeta, None, None, None, None, None, None


class _ConvBlock(torch.autograd.Function):
    """QConv2D -> IQBN(batch statistics) -> act as ONE autograd node (the reference's `Conv.forward`, conv.py:805-809).
    Same kernels as the separate nodes, but the backward knows that the conv is the only consumer of the IQBN input
    gradient: when the conv's backward reads G = M^T dY (separable tensor-core form / direct engine) the IQBN backward
    emits G itself and the mix pass over the gradient disappears."""

    @staticmethod
    def forward(ctx, x, w_r, w_i, w_j, w_k, gamma, beta, running_mean, running_var, stride, padding, dilation, groups, mix,
                algo, eps, momentum, act):
        x, layout = ops.as_layout(x, _state["layout"])
        ws = (w_r, w_i, w_j, w_k)
        g32, b32 = ops._f32c(gamma), ops._f32c(beta)
        # one C call: conv (its tensor-core epilogue may already leave the partial IQBN sums) -> statistics -> apply + act
        y, out, stats = ops.conv_block_fwd(x, ws, g32, b32, running_mean, running_var, stride, padding, dilation, groups, mix,
                                           algo, eps, momentum, act, layout, _state["epilogue_stats"])
        ctx.save_for_backward(x, y, stats, g32, b32, w_r, w_i, w_j, w_k)
        ctx.conf = (tuple(stride), tuple(padding), tuple(dilation), int(groups), tuple(mix), algo, layout, act, gamma.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, y, stats, g32, b32, w_r, w_i, w_j, w_k = ctx.saved_tensors
        stride, padding, dilation, groups, mix, algo, layout, act, pdtype = ctx.conf
        ws = (w_r, w_i, w_j, w_k)
        dout, _ = ops.as_layout(dout, layout)
        if dout.dtype != y.dtype:
            dout = dout.to(y.dtype)
        need_dx = ctx.needs_input_grad[0]
        need_dw = any(ctx.needs_input_grad[1:5])
        # one C call: IQBN backward sums -> apply (emitting G = M^T dY when the conv backward reads G) -> dgrad / wgrad
        dx, dws, dgamma, dbeta = ops.conv_block_bwd(dout, x, y, ws, stats, g32, b32, stride, padding, dilation, groups, mix,
                                                    algo, act, layout, need_dx, need_dw)
        if dws is None:
            dws = [None] * 4
        else:
            dws = [g.to(w.dtype) if g.dtype != w.dtype else g for g, w in zip(dws, ws)]
        if pdtype != torch.float32:
            dgamma, dbeta = dgamma.to(pdtype), dbeta.to(pdtype)
        return (dx, dws[0], dws[1], dws[2], dws[3], dgamma, dbeta) + (None,) * 11


def conv_iqbn_act(x, w_r, w_i, w_j, w_k, gamma, beta, running_mean, running_var, stride, padding, dilation, groups, mix,
                  algo=ALGO_AUTO, eps: float = 1e-5, momentum: float = 0.1, act: int = ACT_SILU) -> torch.Tensor:
    """Training-mode `Conv` block (bias-free QConv2D -> IQBN with batch statistics -> act) as one autograd node."""
    return _ConvBlock.apply(x, w_r, w_i, w_j, w_k, gamma, beta, running_mean, running_var, stride, padding, dilation, groups,
                            mix, algo, eps, momentum, act)


class _IQBNEval(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, eps, act):
        x, layout = ops.as_layout(x)
        g32, b32 = ops._f32c(gamma), ops._f32c(beta)
        rm, rv = ops._f32c(running_mean), ops._f32c(running_var)
        ctx.save_for_backward(x, g32, b32, rm, rv)
        ctx.conf = (layout, act, eps)
        stats = ops.iqbn_eval_stats(g32, b32, rm, rv, eps)          # [20C] table once, then the streaming apply kernel
        return ops.iqbn_apply_fwd(x, layout, stats, g32, b32, act)

    @staticmethod
    def backward(ctx, dy):
        x, g32, b32, rm, rv = ctx.saved_tensors
        layout, act, eps = ctx.conf
        dy, _ = ops.as_layout(dy, layout)
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        dx = ops.iqbn_eval_bwd(dy, x, layout, g32, b32, rm, rv, eps, act) if ctx.needs_input_grad[0] else None
        dgamma = dbeta = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            # frozen-statistics fine-tuning: the reference's eval branch is plain autograd (conv.py:546-552) and does give
            # dgamma = sum dz * xhat, dbeta = sum dz — the two sums of the training backward, taken with the running statistics
            C_ = x.size(1)
            sums = ops.iqbn_bwd_reduce(dy, x, layout, ops.iqbn_eval_stats(g32, b32, rm, rv, eps), g32, b32, act, 0.0)
            dbeta = sums[:4 * C_].to(torch.float32).view(C_, 4)
            dgamma = sums[4 * C_:8 * C_].to(torch.float32).view(C_, 4)
        return dx, dgamma, dbeta, None, None, None, None


def iqbn(x: torch.Tensor, gamma, beta, running_mean, running_var, training: bool, eps: float = 1e-5,
         momentum: float = 0.1, act: int = ACT_NONE, process_group=None) -> torch.Tensor:
    """IQBN on a BCHWQ tensor (conv.py:520-571); `act=ACT_SILU` fuses the SiLU that follows it in `Conv`."""
    if training:
        return _IQBNTrain.apply(x, gamma, beta, running_mean, running_var, eps, momentum, act, process_group)
    return _IQBNEval.apply(x, gamma, beta, running_mean, running_var, eps, act)


class _QUpsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return ops.qupsample_fwd(x, scale)

    @staticmethod
    def backward(ctx, dy):
        return ops.qupsample_bwd(dy, ctx.scale), None


def qupsample_nearest(x: torch.Tensor, scale: int = 2) -> torch.Tensor:
    return _QUpsample.apply(x, int(scale))


class _QAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, heads, key_dim, head_dim, scale):
        qkv, _ = ops.as_layout(qkv, ops.LAYOUT_BHWQC)
        o, lse = ops.qattention_fwd(qkv, heads, key_dim, head_dim, scale)
        ctx.save_for_backward(qkv, o, lse)
        ctx.conf = (heads, key_dim, head_dim, scale)
        return o

    @staticmethod
    def backward(ctx, d_o):
        qkv, o, lse = ctx.saved_tensors
        return ops.qattention_bwd(qkv, o, d_o, lse, *ctx.conf), None, None, None, None


def qattention(qkv: torch.Tensor, heads: int, key_dim: int, head_dim: int, scale: float) -> torch.Tensor:
    """The attention arithmetic of QAttention.forward (block.py:1520-1540) on the qkv projection, fused: per quaternion component
    and head, softmax(q k^T * scale) v over the H*W tokens.  Returns [B, heads*head_dim, H, W, 4] in the internal layout."""
    return _QAttention.apply(qkv, int(heads), int(key_dim), int(head_dim), float(scale))


class _QER(torch.autograd.Function):
    """QER.forward (head.py:40-47) on the tensor-core layout: quan_qer_fwd / quan_qer_bwd."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x, _ = ops.as_layout(x, ops.LAYOUT_BHWQC)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return ops.qer_fwd(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        dx, dw, db = ops.qer_bwd(dy, x, weight, ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2])
        if dw is not None and dw.dtype != weight.dtype:
            dw = dw.to(weight.dtype)
        return dx, None if dw is None else dw.view_as(weight), db


def qer(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """Quaternion -> real 1x1 projection of a [B,C,H,W,4] activation: logical [B,N,H,W], channels-last memory."""
    return _QER.apply(x, weight, bias)


class _QERCat(torch.autograd.Function):
    """`torch.cat((qer_a(xa), qer_b(xb)), 1)` (head.py:143: box and class extractions of one pyramid level) with both extractions
    writing their columns of the concatenated tensor directly, and reading their slices of its gradient in place."""

    @staticmethod
    def forward(ctx, xa, wa, ba, xb, wb, bb):
        xa, _ = ops.as_layout(xa, ops.LAYOUT_BHWQC)
        xb, _ = ops.as_layout(xb, ops.LAYOUT_BHWQC)
        B, _, H, W, _ = xa.shape
        na, nb = wa.size(0), wb.size(0)
        ld = (na + nb + 7) // 8 * 8                    # rows padded to 16 bytes (bf16): vector stores here, vector loads in backward
        buf = torch.empty((B, H, W, ld), dtype=xa.dtype, device=xa.device)
        ops.qer_fwd(xa, wa, ba, buf, 0, na)
        ops.qer_fwd(xb, wb, bb, buf, na, ld - na)      # the padding columns belong to the second extraction (zero-filled)
        ctx.save_for_backward(xa, wa, xb, wb)
        ctx.has_bias = (ba is not None, bb is not None)
        return buf[..., :na + nb].permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xa, wa, xb, wb = ctx.saved_tensors
        na = wa.size(0)
        if dy.dtype != xa.dtype:
            dy = dy.to(xa.dtype)
        dy = ops.pixel_rows(dy)                        # once for both halves; the halves are channel slices of the same rows
        need = ctx.needs_input_grad
        ld = dy.stride(3)                               # rows of the fused head tensor's gradient: na + nb values (+ finite padding)
        pad_ok = ld - na >= (wb.size(0) + 7) // 8 * 8        # the row padding behind the second slice may be read (the kernel clears it)
        dxa, dwa, dba = ops.qer_bwd(dy[:, :na], xa, wa, need[0], need[1], ctx.has_bias[0] and need[2])
        dxb, dwb, dbb = ops.qer_bwd(dy[:, na:], xb, wb, need[3], need[4], ctx.has_bias[1] and need[5], (ld - na) if pad_ok else 0)
        fix = lambda g, w: None if g is None else (g.to(w.dtype) if g.dtype != w.dtype else g).view_as(w)
        return dxa, fix(dwa, wa), dba, dxb, fix(dwb, wb), dbb


def qer_cat(xa, wa, ba, xb, wb, bb) -> torch.Tensor:
    return _QERCat.apply(xa, wa, ba, xb, wb, bb)


class _QCat(torch.autograd.Function):
    """`torch.cat(xs, 1)` of BHWQC tensors / channel chunks (ops.qcat); backward hands every input its channel slice of the gradient — as dense
    tensors written by one launch of quan_rows_split (torch's CatBackward hands out strided views, which every consumer then gathers or
    adds through the generic strided kernel); views when the gradient is stored some other way."""

    @staticmethod
    def forward(ctx, *xs):
        ctx.widths = [x.size(1) for x in xs]
        return ops.qcat(xs)

    @staticmethod
    def backward(ctx, dy):
        import os
        if os.environ.get("QUAN_CAT_SPLIT", "1") != "0":
            parts = ops.qsplit(dy, ctx.widths)         # dense slices, one launch: no gather / strided add downstream
            if parts is not None:
                return tuple(p if ctx.needs_input_grad[i] else None for i, p in enumerate(parts))
        outs, off = [], 0
        for i, c in enumerate(ctx.widths):
            outs.append(dy.narrow(1, off, c) if ctx.needs_input_grad[i] else None)
            off += c
        return tuple(outs)


def qcat(xs) -> torch.Tensor:
    return _QCat.apply(*xs)


class _Chunk2(torch.autograd.Function):
    """`x.chunk(2, 1)` (C2f / C3k2, block.py:350) whose backward re-assembles the two half gradients with one launch of quan_rows_cat;
    torch's SplitBackward builds it from two generic strided copies (~29 us per site at 16 x 128^2)."""

    @staticmethod
    def forward(ctx, x):
        a, b = x.chunk(2, 1)
        ctx.shape = (a.shape, b.shape, x.dtype, x.device)
        return a, b

    @staticmethod
    def backward(ctx, da, db):
        sa, sb, dtype, dev = ctx.shape
        if da is None:
            da = torch.zeros(sa, dtype=dtype, device=dev).contiguous(memory_format=torch.channels_last_3d)
        if db is None:
            db = torch.zeros(sb, dtype=dtype, device=dev).contiguous(memory_format=torch.channels_last_3d)
        if da.is_cuda and da.dim() == 5 and ops.qcat_supported([da, db]):
            return ops.qcat([da, db])
        return torch.cat((da, db), 1)


def chunk2(x: torch.Tensor):
    """The two channel halves of a quaternion activation as views (see _Chunk2); plain `chunk` for anything else."""
    if x.dim() == 5 and x.is_cuda and x.size(1) % 2 == 0 and x.requires_grad and torch.is_grad_enabled():
        return _Chunk2.apply(x)
    return x.chunk(2, 1)


def cat(tensors, dim=0, *, out=None):
    """Drop-in for `torch.cat` inside the reference's block modules (install.py binds it there): channel concatenations of quaternion
    activations in the tensor-core layout take the library's one-launch copy, everything else is torch.cat itself."""
    if out is None and dim == 1 and isinstance(tensors, (list, tuple)) and len(tensors) >= 2 and all(isinstance(t, torch.Tensor) for t in tensors) \
            and tensors[0].dim() == 5 and tensors[0].is_cuda and ops.qcat_supported(tensors):
        return _QCat.apply(*tensors)
    return torch.cat(tensors, dim) if out is None else torch.cat(tensors, dim, out=out)


class _QMaxPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, kernel, stride, padding):
        y, idx = ops.qmaxpool_fwd(x, kernel, stride, padding, with_idx=True)
        ctx.save_for_backward(idx)
        ctx.cfg = (tuple(x.shape[2:4]), kernel, stride, padding)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        in_hw, kernel, stride, padding = ctx.cfg
        return ops.qmaxpool_bwd(dy, idx, in_hw, kernel, stride, padding), None, None, None


def qmaxpool(x: torch.Tensor, kernel_size=2, stride=2, padding=0) -> torch.Tensor:
    """QuaternionMaxPool (block.py:85-109): max pooling of every quaternion component; first maximum wins ties."""
    if not (x.requires_grad and torch.is_grad_enabled()):
        return ops.qmaxpool_fwd(x, kernel_size, stride, padding, with_idx=False)[0]
    return _QMaxPool.apply(x, kernel_size, stride, padding)

