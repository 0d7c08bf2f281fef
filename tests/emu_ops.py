"""TEST INFRASTRUCTURE ONLY: CPU emulation of the `quan_ultralytics_b200.ops` entry points (same signatures, same output shapes,
dtypes and MEMORY FORMATS) written with plain torch ops that restate the reference arithmetic (conv.py:472-499, :553-571,
:388-397, :1229-1246, block.py:85-109).  The product never imports this file and has no CPU path; tests monkeypatch it over
`ops` so that the host logic above the C ABI — autograd wiring, layout propagation through the reference's own Python blocks
(views / reshapes / cats on channels_last_3d tensors), class swap, state dicts, the synced-IQBN protocol — runs in the build
container, which has no GPU.  GPU tests (`-m gpu`) run the same graphs on the real kernels.
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

from quan_ultralytics_b200 import ops

L_BCHWQ, L_BHWQC = ops.LAYOUT_BCHWQ, ops.LAYOUT_BHWQC


def _fmt(t, layout):
    return t.contiguous(memory_format=torch.channels_last_3d if layout == L_BHWQC else torch.contiguous_format)


def _mixmat(m, ref):
    return torch.tensor(list(m), dtype=ref.dtype).view(4, 4)


def _bc(t, C_):
    return t.reshape(C_, 4)[None, :, None, None, :]


def _silu_grad(z):
    s = torch.sigmoid(z)
    return s * (1 + z * (1 - s))


def as_layout(x, layout=None):
    cur = ops.layout_of(x)
    if cur is not None and layout is not None and x.size(1) == 1:
        cur = layout
    if cur is not None and (layout is None or cur == layout):
        return x, cur
    target = layout if layout is not None else (cur if cur is not None else L_BCHWQ)
    return _fmt(x, target), target


def convert_layout(x, dst):
    return _fmt(x, dst)


def poincare_fwd(rgb, out_dtype=torch.float32):
    rgb = rgb.float()
    n = (rgb * rgb).sum(1)
    den = 1 + n
    return torch.stack([(1 - n) / den, 2 * rgb[:, 0] / den, 2 * rgb[:, 1] / den, 2 * rgb[:, 2] / den], -1).unsqueeze(1).to(out_dtype)


def poincare_bwd(rgb, grad_out):
    with torch.enable_grad():
        r = rgb.detach().float().requires_grad_(True)
        poincare_fwd(r).backward(grad_out.float())
    return r.grad


def mix(x, matrix):
    return _fmt(torch.einsum("pq,bchwq->bchwp", _mixmat(matrix, x), x), ops.layout_of(x) or L_BCHWQ)


def qupsample_fwd(x, scale):
    x, layout = as_layout(x)
    return _fmt(x.repeat_interleave(scale, 2).repeat_interleave(scale, 3), layout)


def qupsample_bwd(dy, scale):
    dy, layout = as_layout(dy)
    B, C_, Ho, Wo, _ = dy.shape
    return _fmt(dy.reshape(B, C_, Ho // scale, scale, Wo // scale, scale, 4).sum((3, 5)), layout)


def _planes(x):
    B, C_, H, W, _ = x.shape
    return x.permute(0, 1, 4, 2, 3).reshape(B, C_ * 4, H, W)


def _unplanes(p, C_):
    B, _, H, W = p.shape
    return p.reshape(B, C_, 4, H, W).permute(0, 1, 3, 4, 2)


def qmaxpool_fwd(x, kernel, stride, padding, with_idx=True):
    x, layout = as_layout(x)
    y, idx = F.max_pool2d(_planes(x).float(), kernel, stride, padding, return_indices=True)
    out = _fmt(_unplanes(y.to(x.dtype), x.size(1)), layout)
    # the product stores a one-byte tap; the emulation keeps torch's flat index (int64), only its own bwd reads it
    return out, (_fmt(_unplanes(idx, x.size(1)), layout) if with_idx else None)


def qmaxpool_bwd(dy, idx, in_hw, kernel, stride, padding):
    layout = ops.layout_of(idx)
    C_ = dy.size(1)
    g = F.max_unpool2d(_planes(dy).float().contiguous(), _planes(idx).contiguous(), kernel, stride, padding, output_size=list(in_hw))
    # max_unpool2d writes instead of adding when windows overlap: accumulate explicitly
    B = dy.size(0)
    acc = torch.zeros(B, C_ * 4, in_hw[0] * in_hw[1], dtype=torch.float32)
    acc.scatter_add_(2, _planes(idx).reshape(B, C_ * 4, -1), _planes(dy).float().reshape(B, C_ * 4, -1))
    del g
    return _fmt(_unplanes(acc.view(B, C_ * 4, *in_hw).to(dy.dtype), C_), layout)


# ---- IQBN: stats[20C] = [mean | var(+1e-8) | rstd | 8C unused]; sums[14C] = [sum dz | sum dz*xhat | 6C unused] ---------------
def iqbn_partial_sums(x, layout):
    a = x.double()
    return torch.cat([a.sum((0, 2, 3)).reshape(-1), (a * a).sum((0, 2, 3)).reshape(-1)])


def iqbn_finalize_stats(sums, count, C_, gamma, beta, eps, momentum, rm, rv):
    mean = sums[:4 * C_] / count
    var = sums[4 * C_:8 * C_] / count - mean * mean + 1e-8
    rstd = 1.0 / torch.sqrt(var + eps)
    if rm is not None:
        rm.mul_(1 - momentum).add_(momentum * mean.view(C_, 4).to(rm.dtype))
        rv.mul_(1 - momentum).add_(momentum * var.view(C_, 4).to(rv.dtype))
    return torch.cat([mean, var, rstd, torch.zeros(8 * C_, dtype=torch.float64)]).float()


def iqbn_train_stats(x, layout, gamma, beta, eps, momentum, rm, rv):
    B, C_, H, W, _ = x.shape
    a = x.double()
    mean = a.mean((0, 2, 3))
    var = a.var((0, 2, 3), unbiased=False) + 1e-8
    rstd = 1.0 / torch.sqrt(var + eps)
    if rm is not None:
        rm.mul_(1 - momentum).add_(momentum * mean.to(rm.dtype))
        rv.mul_(1 - momentum).add_(momentum * var.to(rv.dtype))
    return torch.cat([mean.reshape(-1), var.reshape(-1), rstd.reshape(-1), torch.zeros(8 * C_, dtype=torch.float64)]).float()


def iqbn_finalize_partials(*a, **k):
    raise AssertionError("emulation never leaves epilogue partials")


def iqbn_eval_stats(gamma, beta, rm, rv, eps):
    C_ = gamma.size(0)
    return torch.cat([rm.reshape(-1).double(), rv.reshape(-1).double(), 1.0 / torch.sqrt(rv.reshape(-1).double() + eps),
                      torch.zeros(8 * C_, dtype=torch.float64)]).float()


def _ms(stats, C_):
    s = stats.double()
    return s[:4 * C_], s[8 * C_:12 * C_]


def iqbn_apply_fwd(x, layout, stats, gamma, beta, act):
    C_ = x.size(1)
    mean, rstd = _ms(stats, C_)
    z = (x.double() - _bc(mean, C_)) * _bc(rstd, C_) * _bc(gamma.double(), C_) + _bc(beta.double(), C_)
    return _fmt((F.silu(z) if act else z).to(x.dtype), layout)


def iqbn_eval_fwd(x, layout, gamma, beta, rm, rv, eps, act):
    return iqbn_apply_fwd(x, layout, iqbn_eval_stats(gamma, beta, rm, rv, eps), gamma, beta, act)


def _dz(dy, x, stats, gamma, beta, act):
    C_ = x.size(1)
    mean, rstd = _ms(stats, C_)
    xhat = (x.double() - _bc(mean, C_)) * _bc(rstd, C_)
    dz = dy.double()
    if act:
        dz = dz * _silu_grad(xhat * _bc(gamma.double(), C_) + _bc(beta.double(), C_))
    return dz, xhat, rstd


def iqbn_bwd_reduce(dy, x, layout, stats, gamma, beta, act, count=0.0):
    C_ = x.size(1)
    dz, xhat, _ = _dz(dy, x, stats, gamma, beta, act)
    out = torch.zeros(14 * C_, dtype=torch.float64)
    out[:4 * C_] = dz.sum((0, 2, 3)).reshape(-1)
    out[4 * C_:8 * C_] = (dz * xhat).sum((0, 2, 3)).reshape(-1)
    return out


def iqbn_bwd_coef(sums, count, stats, gamma):
    return None


def iqbn_bwd_apply(dy, x, layout, stats, gamma, beta, act, sums, count, want_param_grads=True, mix_t=None):
    C_ = x.size(1)
    dz, xhat, rstd = _dz(dy, x, stats, gamma, beta, act)
    sdz, sdzx = sums[:4 * C_], sums[4 * C_:8 * C_]
    dx = _bc(gamma.double().reshape(-1) * rstd, C_) * (dz - _bc(sdz, C_) / count - xhat * _bc(sdzx, C_) / count)
    if mix_t is not None:
        dx = torch.einsum("pq,bchwq->bchwp", _mixmat(mix_t, dx), dx)
    dg = sdzx.view(C_, 4).float() if want_param_grads else None
    db = sdz.view(C_, 4).float() if want_param_grads else None
    return _fmt(dx.to(x.dtype), layout), dg, db


def iqbn_eval_bwd(dy, x, layout, gamma, beta, rm, rv, eps, act):
    C_ = x.size(1)
    stats = iqbn_eval_stats(gamma, beta, rm, rv, eps)
    dz, _, rstd = _dz(dy, x, stats, gamma, beta, act)
    return _fmt((dz * _bc(gamma.double().reshape(-1) * rstd, C_)).to(x.dtype), layout)


# ---- QConv2D --------------------------------------------------------------------------------------------------------------------
def _conv_core(x, ws, bias_r, stride, padding, dilation, groups, M):
    S = [F.conv2d(x[..., q], ws[q].to(x.dtype), bias_r.to(x.dtype) if (q == 0 and bias_r is not None) else None, stride, padding,
                  dilation, groups) for q in range(4)]
    return torch.stack([sum(M[p, q] * S[q] for q in range(4)) for p in range(4)], -1)


def qconv2d_fwd(x, weights, bias_r, stride, padding, dilation, groups, mix_matrix, algo=0, layout=None, with_stats=False):
    x, layout = as_layout(x, layout)
    xx = x.float() if x.dtype == torch.bfloat16 else x
    y = _conv_core(xx, [w.detach() for w in weights], None if bias_r is None else bias_r.detach(), tuple(stride), tuple(padding),
                   tuple(dilation), groups, _mixmat(mix_matrix, xx))
    y = _fmt(y.to(x.dtype), layout)
    return (y, 0) if with_stats else y


def qconv2d_bwd(dy, x, weights, stride, padding, dilation, groups, mix_matrix, need_dx=True, need_dw=True, need_db=False, algo=0,
                premixed=False):
    dy, layout = as_layout(dy)
    x, _ = as_layout(x, layout)
    cd = torch.float32 if x.dtype == torch.bfloat16 else x.dtype
    M = _mixmat(mix_matrix, torch.empty(0, dtype=cd))
    if premixed:                       # dy holds G = M^T dY: undo the mix so autograd below re-applies it (M is invertible)
        dyy = torch.einsum("pq,bchwq->bchwp", torch.linalg.inv(M.double().T).to(cd), dy.to(cd))
    else:
        dyy = dy.to(cd)
    with torch.enable_grad():
        xg = x.detach().to(cd).requires_grad_(True)
        wg = [w.detach().to(cd).requires_grad_(True) for w in weights]
        bg = torch.zeros(weights[0].size(0), dtype=cd, requires_grad=True)
        y = _conv_core(xg, wg, bg, tuple(stride), tuple(padding), tuple(dilation), groups, M)
        gx, *gw, gb = torch.autograd.grad(y, [xg, *wg, bg], dyy)
    dx = _fmt(gx.to(x.dtype), layout) if need_dx else None
    dws = [g.float() for g in gw] if need_dw else None
    return dx, dws, (gb.float() if need_db else None)


def qconv2d_bwd_wants_mixed(*a, **k):
    return False


def conv_block_fwd(x, weights, gamma, beta, rm, rv, stride, padding, dilation, groups, mix_matrix, algo, eps, momentum, act, layout,
                   epilogue_stats=True):
    y = qconv2d_fwd(x, weights, None, stride, padding, dilation, groups, mix_matrix, algo, layout)
    stats = iqbn_train_stats(y, layout, gamma, beta, eps, momentum, rm, rv)
    return y, iqbn_apply_fwd(y, layout, stats, gamma, beta, act), stats


def conv_block_bwd(dout, x, y, weights, stats, gamma, beta, stride, padding, dilation, groups, mix_matrix, algo, act, layout, need_dx,
                   need_dw):
    B, C_, H, W, _ = y.shape
    cnt = float(B * H * W)
    sums = iqbn_bwd_reduce(dout, y, layout, stats, gamma, beta, act, cnt)
    dy, dg, db = iqbn_bwd_apply(dout, y, layout, stats, gamma, beta, act, sums, cnt)
    dx, dws, _ = qconv2d_bwd(dy, x, weights, stride, padding, dilation, groups, mix_matrix, need_dx, need_dw, False, algo)
    return dx, dws, dg, db


def conv_block_eval_fwd(x, weights, gamma, beta, rm, rv, stride, padding, dilation, groups, mix_matrix, algo, eps, act, layout):
    y = qconv2d_fwd(x, weights, None, stride, padding, dilation, groups, mix_matrix, algo, layout)
    return iqbn_eval_fwd(y, layout, gamma, beta, rm, rv, eps, act)


_NAMES = ["as_layout", "convert_layout", "poincare_fwd", "poincare_bwd", "mix", "qupsample_fwd", "qupsample_bwd", "qmaxpool_fwd",
          "qmaxpool_bwd", "iqbn_partial_sums", "iqbn_finalize_stats", "iqbn_train_stats", "iqbn_finalize_partials", "iqbn_eval_stats",
          "iqbn_apply_fwd", "iqbn_eval_fwd", "iqbn_bwd_reduce", "iqbn_bwd_coef", "iqbn_bwd_apply", "iqbn_eval_bwd", "qconv2d_fwd",
          "qconv2d_bwd", "qconv2d_bwd_wants_mixed", "conv_block_fwd", "conv_block_bwd", "conv_block_eval_fwd"]


@contextlib.contextmanager
def emulated(on_device: bool = True):
    """Patch the emulation over quan_ultralytics_b200.ops for the duration of the block; `on_device` makes `ops.on_device(x)` true for
    CPU tensors so that the fused `Conv` autograd node (functional._ConvBlock) is the path exercised."""
    g = globals()
    saved = {n: getattr(ops, n) for n in _NAMES + ["_require_cuda", "on_device"] if hasattr(ops, n)}
    try:
        for n in _NAMES:
            setattr(ops, n, g[n])
        ops._require_cuda = lambda *ts: None
        ops.on_device = (lambda x: True) if on_device else (lambda x: x.is_cuda)
        yield ops
    finally:
        for n, v in saved.items():
            setattr(ops, n, v)


# ---- QAttention core (block.py:1520-1540) ----------------------------------------------------------------------------------------
def _qattn_ref(qkv, heads, K, V, scale):
    B, Cq, H, W, Q = qkv.shape
    N = H * W
    q, k, v = torch.split(qkv, [heads * K, heads * K, heads * V], dim=1)
    q = q.reshape(B, heads, K, N, Q).permute(0, 1, 4, 3, 2)
    k = k.reshape(B, heads, K, N, Q).permute(0, 1, 4, 2, 3)
    v = v.reshape(B, heads, V, N, Q).permute(0, 1, 4, 3, 2)
    attn = (torch.matmul(q, k) * scale).softmax(dim=-1)
    return torch.matmul(attn, v).permute(0, 1, 4, 3, 2).reshape(B, heads * V, H, W, Q)


def qattention_fwd(qkv, heads, key_dim, head_dim, scale):
    o = _qattn_ref(qkv.float(), heads, key_dim, head_dim, scale).to(qkv.dtype)
    return _fmt(o, L_BHWQC), torch.zeros(1)


def qattention_bwd(qkv, o, d_o, lse, heads, key_dim, head_dim, scale):
    with torch.enable_grad():
        x = qkv.detach().float().requires_grad_(True)
        _qattn_ref(x, heads, key_dim, head_dim, scale).backward(d_o.float())
    return _fmt(x.grad.to(qkv.dtype), L_BHWQC)


_NAMES += ["qattention_fwd", "qattention_bwd"]


# ---- QER (head.py:40-47) -------------------------------------------------------------------------------------------------------
def _qer_ref(x, weight, bias):
    B, C_, H, W, Q = x.shape
    return F.conv2d(x.permute(0, 1, 4, 2, 3).reshape(B, C_ * Q, H, W), weight, bias)


def qer_fwd(x, weight, bias, out=None, col0=0):
    y = _qer_ref(x.float(), weight.float(), None if bias is None else bias.float()).to(x.dtype)
    if out is None:
        return y.contiguous(memory_format=torch.channels_last)
    out[..., col0:col0 + y.shape[1]] = y.permute(0, 2, 3, 1)
    return out


def qer_bwd(dy, x, weight, need_dx, need_dw, need_db):
    with torch.enable_grad():
        xx = x.detach().float().requires_grad_(True)
        ww = weight.detach().float().requires_grad_(True)
        bb = torch.zeros(weight.size(0), requires_grad=True)
        _qer_ref(xx, ww, bb).backward(dy.float())
    dx = _fmt(xx.grad.to(x.dtype), L_BHWQC) if need_dx else None
    return dx, (ww.grad if need_dw else None), (bb.grad if need_db else None)


_NAMES += ["qer_fwd", "qer_bwd"]
