"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  A numpy (float64-capable) restatement of the reference's quaternion
layer arithmetic.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this; the product (quan_ultralytics_b200/) never does and has no CPU fallback.

Parity pinning: the reference ships NO golden vectors or tests for this path, so this oracle
is pinned against outputs of the reference itself, generated in the build container by importing
/root/reference (tests/golden/make_golden.py -> tests/golden/*.npz) and checked in tests/test_oracle_golden.py.

Each function cites the reference lines it restates (paths relative to the reference repo root).  Third-party
arithmetic the reference delegates to — torch.nn.functional.conv2d, Tensor.mean/var, F.interpolate (torch,
unpinned: pyproject.toml:71 `torch>=1.8.0`) — is restated from its published definition (cross-correlation with
zero padding; biased variance; nearest-neighbour index floor(i/scale)).
"""
from __future__ import annotations

import numpy as np

# ultralytics/nn/modules/conv.py:493-496 (the PyTorch path of the YOLO models)
M_A = np.array([[1, -1, -1, -1], [-1, 1, 1, -1], [-1, -1, 1, 1], [-1, 1, -1, 1]], dtype=np.float64)
# classification/quaternion/qconv.py:606-609 == ultralytics/nn/cuda/quaternion_ops.cu:152-155
M_B = np.array([[1, 1, 1, 1], [1, -1, -1, 1], [1, 1, -1, -1], [1, -1, 1, -1]], dtype=np.float64)
MIX = {"A": M_A, "B": M_B}


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def conv_out_size(n, k, s, p, d):
    return (n + 2 * p - d * (k - 1) - 1) // s + 1


# ---------------------------------------------------------------------------------------------------------------
# Poincare map — conv.py:388-397 (identical at classification/quaternion/qconv.py:524-533)
# ---------------------------------------------------------------------------------------------------------------
def poincare_fwd(rgb: np.ndarray) -> np.ndarray:
    """rgb [B,3,H,W] -> [B,1,H,W,4]: n=|x|^2, q=[(1-n)/(1+n), 2x/(1+n)]."""
    n = np.sum(rgb * rgb, axis=1)                       # conv.py:393
    den = 1.0 + n                                       # :394
    real = (1.0 - n) / den                              # :395
    vec = 2.0 * rgb / den[:, None]                      # :396
    out = np.stack([real, vec[:, 0], vec[:, 1], vec[:, 2]], axis=-1)   # :397
    return out[:, None]                                 # :408 unsqueeze(1)


def poincare_bwd(rgb: np.ndarray, gout: np.ndarray) -> np.ndarray:
    """Analytic VJP of poincare_fwd (the reference uses autograd)."""
    g = gout[:, 0]                                      # [B,H,W,4]
    n = np.sum(rgb * rgb, axis=1)
    inv = 1.0 / (1.0 + n)
    dot = rgb[:, 0] * g[..., 1] + rgb[:, 1] * g[..., 2] + rgb[:, 2] * g[..., 3]
    common = -4.0 * inv * inv * (g[..., 0] + dot)
    return np.stack([common * rgb[:, a] + 2.0 * inv * g[..., a + 1] for a in range(3)], axis=1)


# ---------------------------------------------------------------------------------------------------------------
# QConv2D — conv.py:472-499 (M_A), classification/quaternion/qconv.py:592-612 (M_B)
# ---------------------------------------------------------------------------------------------------------------
def _im2col(xq: np.ndarray, kH, kW, sH, sW, pH, pW, dH, dW):
    """xq [B,C,H,W] -> cols [B,C,kH,kW,Ho,Wo] (zero padded cross-correlation windows; F.conv2d semantics)."""
    B, C, H, W = xq.shape
    Ho, Wo = conv_out_size(H, kH, sH, pH, dH), conv_out_size(W, kW, sW, pW, dW)
    xp = np.zeros((B, C, H + 2 * pH, W + 2 * pW), dtype=xq.dtype)
    xp[:, :, pH:pH + H, pW:pW + W] = xq
    cols = np.empty((B, C, kH, kW, Ho, Wo), dtype=xq.dtype)
    for kh in range(kH):
        for kw in range(kW):
            h0, w0 = kh * dH, kw * dW
            cols[:, :, kh, kw] = xp[:, :, h0:h0 + (Ho - 1) * sH + 1:sH, w0:w0 + (Wo - 1) * sW + 1:sW]
    return cols


def _conv2d(xq, wq, stride, padding, dilation, groups):
    """Real grouped conv2d of one component: xq [B,Ci,H,W], wq [Co,Ci/g,kH,kW] -> [B,Co,Ho,Wo]."""
    (sH, sW), (pH, pW), (dH, dW) = stride, padding, dilation
    Co, Cig, kH, kW = wq.shape
    cols = _im2col(xq, kH, kW, sH, sW, pH, pW, dH, dW)
    B = xq.shape[0]
    Cog = Co // groups
    outs = []
    for g in range(groups):
        outs.append(np.einsum("bcklhw,ockl->bohw", cols[:, g * Cig:(g + 1) * Cig], wq[g * Cog:(g + 1) * Cog],
                              optimize=True))
    return np.concatenate(outs, axis=1)


def qconv2d_fwd(x, w, bias_r=None, stride=1, padding=0, dilation=1, groups=1, mix=M_A):
    """x [B,Ci,H,W,4]; w = 4 arrays [Co,Ci/g,kH,kW]; returns y [B,Co,Ho,Wo,4].
    S_q = conv2d(x[...,q], w_q) (conv.py:480-483; bias_r joins S_r only), y_p = sum_q M[p,q] S_q (:485-499)."""
    stride, padding, dilation = _pair(stride), _pair(padding), _pair(dilation)
    S = [_conv2d(x[..., q], w[q], stride, padding, dilation, groups) for q in range(4)]
    if bias_r is not None:
        S[0] = S[0] + bias_r[None, :, None, None]
    S = np.stack(S, axis=-1)                             # [B,Co,Ho,Wo,4(q)]
    return np.einsum("pq,bchwq->bchwp", np.asarray(mix, dtype=S.dtype), S)


def qconv2d_bwd(dy, x, w, stride=1, padding=0, dilation=1, groups=1, mix=M_A, has_bias=False):
    """Analytic backward of qconv2d_fwd: G = M^T dY; dX_q = conv_transpose(G_q, w_q); dW_q = corr(G_q, x_q);
    db_r = sum G_r.  (What torch.autograd derives for conv.py:472-499; the reference extension's own kernels are
    ultralytics/nn/cuda/quaternion_ops.cu:185-530.)  Returns (dx, [dw_q]*4, db_r or None)."""
    (sH, sW), (pH, pW), (dH, dW) = _pair(stride), _pair(padding), _pair(dilation)
    B, Ci, H, W, _ = x.shape
    Co, Cig, kH, kW = w[0].shape
    Cog = Co // groups
    G = np.einsum("pq,bchwp->bchwq", np.asarray(mix, dtype=dy.dtype), dy)     # M^T dY
    Ho, Wo = G.shape[2], G.shape[3]
    dx = np.zeros_like(x)
    dws = []
    for q in range(4):
        cols = _im2col(x[..., q], kH, kW, sH, sW, pH, pW, dH, dW)           # [B,Ci,kH,kW,Ho,Wo]
        dw = np.empty_like(w[q])
        dxp = np.zeros((B, Ci, H + 2 * pH, W + 2 * pW), dtype=x.dtype)
        for g in range(groups):
            Gg = G[:, g * Cog:(g + 1) * Cog, :, :, q]                        # [B,Cog,Ho,Wo]
            dw[g * Cog:(g + 1) * Cog] = np.einsum("bohw,bcklhw->ockl", Gg, cols[:, g * Cig:(g + 1) * Cig],
                                                  optimize=True)
            dcols = np.einsum("bohw,ockl->bcklhw", Gg, w[q][g * Cog:(g + 1) * Cog], optimize=True)
            for kh in range(kH):
                for kw in range(kW):
                    h0, w0 = kh * dH, kw * dW
                    dxp[:, g * Cig:(g + 1) * Cig, h0:h0 + (Ho - 1) * sH + 1:sH, w0:w0 + (Wo - 1) * sW + 1:sW] += \
                        dcols[:, :, kh, kw]
        dx[..., q] = dxp[:, :, pH:pH + H, pW:pW + W]
        dws.append(dw)
    db = G[..., 0].sum(axis=(0, 2, 3)) if has_bias else None
    return dx, dws, db


# ---------------------------------------------------------------------------------------------------------------
# IQBN — conv.py:553-571 (train), :546-552 (eval); classification/quaternion/qconv.py:378-396
# ---------------------------------------------------------------------------------------------------------------
def silu(z):
    return z / (1.0 + np.exp(-z))


def silu_grad(z):
    s = 1.0 / (1.0 + np.exp(-z))
    return s * (1.0 + z * (1.0 - s))


def iqbn_train_fwd(x, gamma, beta, running_mean=None, running_var=None, eps=1e-5, momentum=0.1, act=False):
    """Returns (y, new_running_mean, new_running_var, (mean, var, rstd)).  x [B,C,H,W,4]; params [C,4]."""
    mean = x.mean(axis=(0, 2, 3))                        # conv.py:556  [C,4]
    var = x.var(axis=(0, 2, 3)) + 1e-8                   # :557 (unbiased=False) + 1e-8
    new_rm = new_rv = None
    if running_mean is not None:
        new_rm = (1 - momentum) * running_mean + momentum * mean     # :561
        new_rv = (1 - momentum) * running_var + momentum * var       # :562
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mean[None, :, None, None, :]) * rstd[None, :, None, None, :]   # :566
    z = xhat * gamma[None, :, None, None, :] + beta[None, :, None, None, :]    # :569-571
    y = silu(z) if act else z                            # conv.py:809 nn.SiLU after bn
    return y, new_rm, new_rv, (mean, var, rstd)


def iqbn_train_bwd(dy, x, gamma, beta, eps=1e-5, act=False):
    """Analytic backward of iqbn_train_fwd (reference: autograd).  Returns (dx, dgamma, dbeta)."""
    mean = x.mean(axis=(0, 2, 3))
    var = x.var(axis=(0, 2, 3)) + 1e-8
    rstd = 1.0 / np.sqrt(var + eps)
    bc = lambda a: a[None, :, None, None, :]
    xhat = (x - bc(mean)) * bc(rstd)
    dz = dy * silu_grad(xhat * bc(gamma) + bc(beta)) if act else dy
    n = x.shape[0] * x.shape[2] * x.shape[3]
    dbeta = dz.sum(axis=(0, 2, 3))
    dgamma = (dz * xhat).sum(axis=(0, 2, 3))
    dx = bc(gamma * rstd) * (dz - bc(dbeta) / n - xhat * bc(dgamma) / n)
    return dx, dgamma, dbeta


def iqbn_eval_fwd(x, gamma, beta, running_mean, running_var, eps=1e-5, act=False):
    bc = lambda a: a[None, :, None, None, :]
    z = (x - bc(running_mean)) / np.sqrt(bc(running_var) + eps) * bc(gamma) + bc(beta)   # conv.py:546-552
    return silu(z) if act else z


def iqbn_eval_bwd(dy, x, gamma, beta, running_mean, running_var, eps=1e-5, act=False):
    bc = lambda a: a[None, :, None, None, :]
    scale = bc(gamma) / np.sqrt(bc(running_var) + eps)
    z = (x - bc(running_mean)) * scale + bc(beta)
    dz = dy * silu_grad(z) if act else dy
    return dz * scale


# ---------------------------------------------------------------------------------------------------------------
# QUpsample — conv.py:1229-1246 (F.interpolate(mode='nearest') per component: src = floor(dst / scale))
# ---------------------------------------------------------------------------------------------------------------
def qupsample_fwd(x, scale=2):
    return np.repeat(np.repeat(x, scale, axis=2), scale, axis=3)


def qupsample_bwd(dy, scale=2):
    B, C, Ho, Wo, Q = dy.shape
    return dy.reshape(B, C, Ho // scale, scale, Wo // scale, scale, Q).sum(axis=(3, 5))


# ---------------------------------------------------------------------------------------------------------------
# QuaternionMaxPool — block.py:85-109 (nn.MaxPool2d per component + stack); backward = autograd of it.
# nn.MaxPool2d: padding is -inf, the window is scanned row-major and a later element replaces the running maximum only
# if strictly greater (or NaN): the first maximum wins ties and alone receives the gradient.
# ---------------------------------------------------------------------------------------------------------------
def _pool_windows(H, W, k, s, p):
    (kh, kw), (sh, sw), (ph, pw) = [(v, v) if isinstance(v, int) else tuple(v) for v in (k, s, p)]
    Ho, Wo = (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1
    return kh, kw, sh, sw, ph, pw, Ho, Wo


def qmaxpool_fwd(x, kernel_size, stride, padding):
    """x [B,C,H,W,4] -> (y [B,C,Ho,Wo,4], arg [B,C,Ho,Wo,4] flat input index hi*W + wi of the winner)."""
    B, C, H, W, Q = x.shape
    kh, kw, sh, sw, ph, pw, Ho, Wo = _pool_windows(H, W, kernel_size, stride, padding)
    y = np.full((B, C, Ho, Wo, Q), -np.inf, dtype=x.dtype)
    arg = np.full((B, C, Ho, Wo, Q), -1, dtype=np.int64)
    ho, wo = np.meshgrid(np.arange(Ho), np.arange(Wo), indexing="ij")
    for a in range(kh):
        for b in range(kw):
            hi, wi = ho * sh - ph + a, wo * sw - pw + b
            ok = (hi >= 0) & (hi < H) & (wi >= 0) & (wi < W)
            v = x[:, :, np.clip(hi, 0, H - 1), np.clip(wi, 0, W - 1), :]            # [B,C,Ho,Wo,Q]
            take = ok[None, None, :, :, None] & ((arg < 0) | (v > y) | np.isnan(v))
            y = np.where(take, v, y)
            arg = np.where(take, (hi * W + wi)[None, None, :, :, None], arg)
    return y, arg


def qmaxpool_bwd(dy, arg, in_hw):
    B, C, Ho, Wo, Q = dy.shape
    H, W = in_hw
    dx = np.zeros((B, C, H * W, Q), dtype=dy.dtype)
    bi, ci, qi = np.meshgrid(np.arange(B), np.arange(C), np.arange(Q), indexing="ij")
    for ho in range(Ho):
        for wo in range(Wo):
            np.add.at(dx, (bi, ci, arg[:, :, ho, wo, :], qi), dy[:, :, ho, wo, :])
    return dx.reshape(B, C, H, W, Q)


def mix_apply(x, mix):
    return np.einsum("pq,bchwq->bchwp", np.asarray(mix, dtype=x.dtype), x)


# ---- optimizer step --------------------------------------------------------------------------------------
def sgd_clip_step(params, grads, bufs, groups, lrs, wds, momentum, max_norm, nesterov=True, dampening=0.0):
    """ultralytics/engine/trainer.py:586-594 optimizer_step: torch.nn.utils.clip_grad_norm_(parameters, max_norm) — total =
    ||(||g_1||, ..., ||g_n||)||_2, every gradient scaled by min(1, max_norm / (total + 1e-6)) — then torch.optim.SGD.step():
    g += wd p; buf = momentum buf + (1 - dampening) g (first step: buf = g, which a zero buffer reproduces for dampening = 0);
    step = g + momentum buf if nesterov else buf; p -= lr step.  Lists of fp64 arrays in, (new params, new bufs, clipped grads,
    total norm) out; `groups[i]` selects lrs / wds for tensor i; max_norm <= 0 disables clipping."""
    total = float(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads)))
    k = min(1.0, max_norm / (total + 1e-6)) if max_norm and max_norm > 0 else 1.0
    out_p, out_b, out_g = [], [], []
    for p, g, b, gi in zip(params, grads, bufs, groups):
        gc = g * k
        gg = gc + wds[gi] * p
        nb = momentum * b + (1.0 - dampening) * gg
        step = gg + momentum * nb if nesterov else nb
        out_p.append(p - lrs[gi] * step)
        out_b.append(nb)
        out_g.append(gc)
    return out_p, out_b, out_g, total


def ema_update(ema, value, decay):
    """ultralytics/utils/torch_utils.py:514-525 ModelEMA.update: v *= d; v += (1 - d) * model value."""
    return decay * ema + (1.0 - decay) * value


# ---- QAttention core --------------------------------------------------------------------------------------
def qattention_fwd(qkv, heads, key_dim, head_dim, scale=None):
    """ultralytics/nn/modules/block.py:1520-1540: qkv [B, heads*(2K+V), H, W, 4] -> [B, heads*V, H, W, 4].  split on dim 1 into
    q | k | v (head-major inside each), tokens n = h*W + w, per (b, head, component): softmax(q k^T * K^-0.5) v."""
    B, Cq, H, W, Q = qkv.shape
    N = H * W
    K, V = key_dim, head_dim
    scale = K ** -0.5 if scale is None else scale
    q = qkv[:, :heads * K].reshape(B, heads, K, N, Q)
    k = qkv[:, heads * K:2 * heads * K].reshape(B, heads, K, N, Q)
    v = qkv[:, 2 * heads * K:].reshape(B, heads, V, N, Q)
    s = np.einsum("bhknq,bhkmq->bhqnm", q, k) * scale                 # [B, h, 4, N, N]
    s = s - s.max(axis=-1, keepdims=True)
    a = np.exp(s)
    a /= a.sum(axis=-1, keepdims=True)
    o = np.einsum("bhqnm,bhvmq->bhvnq", a, v)                         # [B, h, V, N, 4]
    return o.reshape(B, heads * V, H, W, Q), a


def qattention_bwd(d_o, qkv, heads, key_dim, head_dim, scale=None):
    """Analytic VJP of qattention_fwd (what autograd computes for block.py:1530-1536)."""
    B, Cq, H, W, Q = qkv.shape
    N = H * W
    K, V = key_dim, head_dim
    scale = K ** -0.5 if scale is None else scale
    _, a = qattention_fwd(qkv, heads, K, V, scale)
    q = qkv[:, :heads * K].reshape(B, heads, K, N, Q)
    k = qkv[:, heads * K:2 * heads * K].reshape(B, heads, K, N, Q)
    v = qkv[:, 2 * heads * K:].reshape(B, heads, V, N, Q)
    do = d_o.reshape(B, heads, V, N, Q)
    dv = np.einsum("bhqnm,bhvnq->bhvmq", a, do)
    da = np.einsum("bhvnq,bhvmq->bhqnm", do, v)
    ds = a * (da - (a * da).sum(axis=-1, keepdims=True)) * scale
    dq = np.einsum("bhqnm,bhkmq->bhknq", ds, k)
    dk = np.einsum("bhqnm,bhknq->bhkmq", ds, q)
    return np.concatenate([dq.reshape(B, heads * K, H, W, Q), dk.reshape(B, heads * K, H, W, Q), dv.reshape(B, heads * V, H, W, Q)], axis=1)


# ---- QER: quaternion -> real extraction of the heads ------------------------------------------------------
def qer_fwd(x, weight, bias=None):
    """ultralytics/nn/modules/head.py:40-47: x [B,C,H,W,4] -> permute(0,1,4,2,3).view(B, 4C, H, W) (channel index c*4 + q) -> the real
    1x1 convolution `output_proj` (head.py:36): out[b,n,h,w] = bias[n] + sum_{c,q} weight[n, c*4+q] x[b,c,h,w,q]."""
    B, C_, H, W, Q = x.shape
    w = np.asarray(weight).reshape(weight.shape[0], C_, Q)
    out = np.einsum("bchwq,ncq->bnhw", x, w)
    return out if bias is None else out + np.asarray(bias)[None, :, None, None]


def qer_bwd(dy, x, weight):
    """VJP of qer_fwd: (dx [B,C,H,W,4], dweight like weight, dbias [N])."""
    B, C_, H, W, Q = x.shape
    w = np.asarray(weight).reshape(weight.shape[0], C_, Q)
    dx = np.einsum("bnhw,ncq->bchwq", dy, w)
    dw = np.einsum("bnhw,bchwq->ncq", dy, x).reshape(weight.shape)
    return dx, dw, dy.sum(axis=(0, 2, 3))
