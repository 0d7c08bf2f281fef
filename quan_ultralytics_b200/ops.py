"""Tensor-level wrappers over the C ABI (include/quan_sm100.h): torch owns memory and streams, the library computes.

Every function here takes CUDA tensors, allocates outputs with torch (so the caching allocator and autograd own the
memory), passes raw pointers + the current stream to libquan_sm100.so and raises RuntimeError on a
non-zero return.  Nothing in this module computes with torch ops.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_SILU, ALGO_AUTO, ALGO_DEPTHWISE, ALGO_DIRECT, ALGO_SMALLC, ALGO_TCGEN05, BF16, F32, LAYOUT_BCHWQ, LAYOUT_BHWQC,
                   ConvDims, PtrArray4, check)

# Mixing matrices of the reference
M_A = (1., -1., -1., -1., -1., 1., 1., -1., -1., -1., 1., 1., -1., 1., -1., 1.)   # ultralytics conv.py:493-496
M_B = (1., 1., 1., 1., 1., -1., -1., 1., 1., 1., -1., -1., 1., -1., 1., -1.)       # classification qconv.py:606-609
MIX = {"A": M_A, "B": M_B}

_FloatArr16 = C.c_float * 16


_mix_cache = {}


def _mix_arg(mix: Sequence[float]):
    """ctypes float[16] of a mixing matrix (cached: the same two or three matrices are passed on every call)."""
    key = tuple(mix)
    arr = _mix_cache.get(key)
    if arr is None:
        assert len(key) == 16
        arr = _mix_cache[key] = _FloatArr16(*[float(v) for v in key])
    return arr


def _mix_t(mix: Sequence[float]) -> Tuple[float, ...]:
    return tuple(mix[p * 4 + q] for q in range(4) for p in range(4))


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise RuntimeError(f"quan ops support float32 and bfloat16 activations, got {t.dtype}")


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("quan ops need CUDA tensors: there is no CPU fallback (oracle/ is test infrastructure only)")


def on_device(x: torch.Tensor) -> bool:
    """True when `x` lives where the library computes (a CUDA device)."""
    return x.is_cuda


def layout_of(x: torch.Tensor) -> Optional[int]:
    """Physical layout code of a logical [B,C,H,W,4] tensor, or None if it is neither supported layout."""
    if x.dim() != 5 or x.size(4) != 4:
        raise RuntimeError(f"expected a BCHWQ tensor [B,C,H,W,4], got {tuple(x.shape)}")
    if x.is_contiguous(memory_format=torch.channels_last_3d) and x.size(1) > 1:
        return LAYOUT_BHWQC
    if x.is_contiguous():
        return LAYOUT_BCHWQ
    if x.is_contiguous(memory_format=torch.channels_last_3d):
        return LAYOUT_BHWQC
    return None


def empty_q(shape, dtype, device, layout: int) -> torch.Tensor:
    fmt = torch.channels_last_3d if layout == LAYOUT_BHWQC else torch.contiguous_format
    return torch.empty(shape, dtype=dtype, device=device, memory_format=fmt)


def _memory_format(layout: int):
    return torch.channels_last_3d if layout == LAYOUT_BHWQC else torch.contiguous_format


def as_layout(x: torch.Tensor, layout: Optional[int] = None) -> Tuple[torch.Tensor, int]:
    """Return (tensor, layout code) with the tensor dense in one of the two layouts (converting only if needed)."""
    cur = layout_of(x)
    if cur is not None and layout is not None and x.size(1) == 1:
        cur = layout          # one quaternion channel (the Poincare output): both layouts are the same bytes
    if cur is not None and (layout is None or cur == layout) and x.data_ptr() % 16 == 0:
        return x, cur
    target = layout if layout is not None else (cur if cur is not None else LAYOUT_BCHWQ)
    if cur is not None and x.data_ptr() % 16 == 0:
        return convert_layout(x, target), target
    if cur is None and layout in (None, LAYOUT_BHWQC):        # a strided channel chunk of a BHWQC tensor: vectorised gather
        g = _gather_channel_slice(x)
        if g is not None:
            return g, LAYOUT_BHWQC
    y = x.contiguous(memory_format=_memory_format(target))
    if y.data_ptr() % 16 != 0:
        y = y.clone(memory_format=_memory_format(target))
    return y, target


def _gather_channel_slice(x: torch.Tensor) -> Optional[torch.Tensor]:
    """A channel chunk of a BHWQC tensor (strides: channel 1, component C_total, then W, H, B multiples of 4*C_total) copied into a
    dense BHWQC tensor by quan_rows_gather; None when `x` is some other kind of view."""
    B, C_, H, W, _ = x.shape
    sb, sc, sh, sw, sq = x.stride()
    if not (x.is_cuda and sc == 1 and sq > C_ and sw == 4 * sq and sh == W * sw and sb == H * sh):
        return None
    esz = x.element_size()
    if (C_ * esz) % 4 or (sq * esz) % 4 or x.data_ptr() % 4:
        return None
    out = torch.empty(x.shape, dtype=x.dtype, device=x.device, memory_format=torch.channels_last_3d)
    check(_lib.load().quan_rows_gather(x.data_ptr(), out.data_ptr(), B * H * W * 4, C_ * esz, sq * esz, _stream(x)), "quan_rows_gather")
    return out


def _row_source(x: torch.Tensor):
    """(pointer, row pitch in bytes, row bytes) when `x` ([B,C,H,W,4]) is stored as one row of C contiguous values per (pixel, component)
    with uniform pitch — a dense BHWQC tensor or a channel chunk of one — else None."""
    B, C_, H, W, _ = x.shape
    sb, sc, sh, sw, sq = x.stride()
    if not ((sc == 1 or C_ == 1) and sq >= C_ and sw == 4 * sq and sh == W * sw and sb == H * sh):
        return None
    esz = x.element_size()
    return x.data_ptr(), sq * esz, C_ * esz


def qcat_supported(tensors) -> bool:
    """True when ops.qcat serves `torch.cat(tensors, 1)`: CUDA [B,C_i,H,W,4] tensors of one float32 / bfloat16 dtype and one spatial shape,
    every one stored in pixel-component rows (dense BHWQC or a channel chunk of it), 4-byte multiples throughout."""
    t0 = tensors[0]
    if len(tensors) < 2 or t0.dtype not in (torch.float32, torch.bfloat16):
        return False
    for t in tensors:
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dim() == 5 and t.size(4) == 4 and t.dtype == t0.dtype and t.device == t0.device
                and t.shape[0] == t0.shape[0] and t.shape[2:] == t0.shape[2:] and t.size(1) > 0):
            return False
        src = _row_source(t)
        if src is None or src[0] % 4 or src[1] % 4 or src[2] % 4:
            return False
    return (sum(t.size(1) for t in tensors) * t0.element_size()) % 4 == 0


def qcat(tensors) -> torch.Tensor:
    """`torch.cat(tensors, 1)` into a dense BHWQC tensor in one launch (quan_rows_cat).  Neighbouring chunks of one tensor (the two
    halves C2f keeps of its first convolution) are moved as one wider source."""
    _require_cuda(*tensors)
    t0 = tensors[0]
    B, _, H, W, _ = t0.shape
    Ct = sum(t.size(1) for t in tensors)
    out = empty_q((B, Ct, H, W, 4), t0.dtype, t0.device, LAYOUT_BHWQC)
    esz = t0.element_size()
    srcs = []
    for t in tensors:
        ptr, ld, rb = _row_source(t)
        if srcs and srcs[-1][1] == ld and srcs[-1][0] + srcs[-1][2] == ptr and ld > srcs[-1][2]:
            srcs[-1] = (srcs[-1][0], ld, srcs[-1][2] + rb)
        else:
            srcs.append((ptr, ld, rb))
    lib = _lib.load()
    nrows = B * H * W * 4
    off = 0
    for i in range(0, len(srcs), _lib.CAT_MAX):                      # more than CAT_MAX sources: several launches
        part = srcs[i:i + _lib.CAT_MAX]
        arr = (_lib.CatSrc * len(part))(*[_lib.CatSrc(p, ld, rb) for p, ld, rb in part])
        check(lib.quan_rows_cat(C.cast(arr, C.c_void_p), len(part), out.data_ptr() + off, Ct * esz, nrows, _stream(t0)), "quan_rows_cat")
        off += sum(rb for _, _, rb in part)
    return out


def qsplit(dy: torch.Tensor, widths) -> list:
    """The gradient of ops.qcat as DENSE BHWQC slices in one launch (quan_rows_split): dy [B, sum(widths), H, W, 4] in pixel-component
    rows (dense, or itself a channel chunk) -> one dense tensor per width.  None when dy is stored some other way."""
    src = _row_source(dy) if (dy.is_cuda and dy.dim() == 5 and dy.dtype in (torch.float32, torch.bfloat16)) else None
    esz = dy.element_size()
    if src is None or src[0] % 4 or src[1] % 4 or any((c * esz) % 4 for c in widths) or len(widths) > _lib.CAT_MAX:
        return None
    B, _, H, W, _ = dy.shape
    outs = [empty_q((B, c, H, W, 4), dy.dtype, dy.device, LAYOUT_BHWQC) for c in widths]
    arr = (_lib.CatSrc * len(outs))(*[_lib.CatSrc(o.data_ptr(), c * esz, c * esz) for o, c in zip(outs, widths)])
    check(_lib.load().quan_rows_split(C.cast(arr, C.c_void_p), len(outs), src[0], src[1], B * H * W * 4, _stream(dy)), "quan_rows_split")
    return outs


# ---- workspaces ------------------------------------------------------------------------------------------------
_ws_cache = {}
_iqbn_ws_cache = {}


def _ws_key(device: torch.device):
    """Workspaces hold live intermediate data between the kernels of one call (G = M^T dY, packed weights, split-K partials, the
    conv epilogue's IQBN partial sums): they are private to the (device, stream) a call runs on, so that side streams, loader
    threads or an overlapped evaluation never share scratch memory with the training stream."""
    return (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    key = _ws_key(device)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _iqbn_workspace(C_: int, device: torch.device) -> torch.Tensor:
    key = _ws_key(device) + (C_,)
    buf = _iqbn_ws_cache.get(key)
    if buf is None:
        n = _lib.load().quan_iqbn_workspace_bytes(C_)
        buf = torch.zeros(n, dtype=torch.uint8, device=device)  # zeroed once; kernels leave it zeroed
        _iqbn_ws_cache[key] = buf
    return buf


def _f32c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        t = t.detach().to(torch.float32).contiguous()
    return t


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


# ---- Poincare ----------------------------------------------------------------------------------------------------
def poincare_fwd(rgb: torch.Tensor, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    _require_cuda(rgb)
    if rgb.dim() != 4 or rgb.size(1) != 3:
        raise RuntimeError(f"poincare_fwd expects [B,3,H,W], got {tuple(rgb.shape)}")
    rgb = _f32c(rgb)
    B, _, H, W = rgb.shape
    out = torch.empty((B, 1, H, W, 4), dtype=out_dtype, device=rgb.device)
    check(_lib.load().quan_poincare_fwd(rgb.data_ptr(), out.data_ptr(), B, H, W, _dtype_code(out), _stream(rgb)),
          "quan_poincare_fwd")
    return out


def poincare_bwd(rgb: torch.Tensor, grad_out: torch.Tensor) -> torch.Tensor:
    _require_cuda(rgb, grad_out)
    rgb = _f32c(rgb)
    B, _, H, W = rgb.shape
    grad_out = grad_out.contiguous()
    grad = torch.empty_like(rgb)
    check(_lib.load().quan_poincare_bwd(rgb.data_ptr(), grad_out.data_ptr(), grad.data_ptr(), B, H, W,
                                        _dtype_code(grad_out), _stream(rgb)), "quan_poincare_bwd")
    return grad


# ---- layout / mix -----------------------------------------------------------------------------------------------
def convert_layout(x: torch.Tensor, dst_layout: int) -> torch.Tensor:
    _require_cuda(x)
    src = layout_of(x)
    if src is None:
        raise RuntimeError("convert_layout: tensor is not dense in either supported layout")
    if src == dst_layout:
        return x
    B, C_, H, W, _ = x.shape
    out = empty_q(x.shape, x.dtype, x.device, dst_layout)
    check(_lib.load().quan_layout_convert(x.data_ptr(), src, out.data_ptr(), dst_layout, B, C_, H, W, _dtype_code(x),
                                          _stream(x)), "quan_layout_convert")
    return out


def mix(x: torch.Tensor, matrix: Sequence[float]) -> torch.Tensor:
    _require_cuda(x)
    x, layout = as_layout(x)
    B, C_, H, W, _ = x.shape
    out = torch.empty_like(x, memory_format=torch.preserve_format)
    check(_lib.load().quan_mix(x.data_ptr(), out.data_ptr(), B, C_, H, W, _dtype_code(x), layout,
                               C.cast(_mix_arg(matrix), C.c_void_p), _stream(x)), "quan_mix")
    return out


# ---- QUpsample ---------------------------------------------------------------------------------------------------
def qupsample_fwd(x: torch.Tensor, scale: int) -> torch.Tensor:
    _require_cuda(x)
    x, layout = as_layout(x)
    B, C_, H, W, _ = x.shape
    out = empty_q((B, C_, H * scale, W * scale, 4), x.dtype, x.device, layout)
    check(_lib.load().quan_qupsample_nearest_fwd(x.data_ptr(), out.data_ptr(), B, C_, H, W, scale, _dtype_code(x),
                                                 layout, _stream(x)), "quan_qupsample_nearest_fwd")
    return out


def qupsample_bwd(dy: torch.Tensor, scale: int) -> torch.Tensor:
    _require_cuda(dy)
    dy, layout = as_layout(dy)
    B, C_, Ho, Wo, _ = dy.shape
    H, W = Ho // scale, Wo // scale
    out = empty_q((B, C_, H, W, 4), dy.dtype, dy.device, layout)
    check(_lib.load().quan_qupsample_nearest_bwd(dy.data_ptr(), out.data_ptr(), B, C_, H, W, scale, _dtype_code(dy),
                                                 layout, _stream(dy)), "quan_qupsample_nearest_bwd")
    return out


# ---- QuaternionMaxPool -------------------------------------------------------------------------------------------
def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def qmaxpool_out_hw(H: int, W: int, kernel, stride, padding) -> Tuple[int, int]:
    (kh, kw), (sh, sw), (ph, pw) = _pair(kernel), _pair(stride), _pair(padding)
    return (H + 2 * ph - kh) // sh + 1, (W + 2 * pw - kw) // sw + 1


def qmaxpool_fwd(x: torch.Tensor, kernel, stride, padding, with_idx: bool = True):
    """nn.MaxPool2d on each quaternion component (block.py:85-109).  Returns (y, idx); idx (uint8, the winning tap per
    output element, laid out like y) is None when with_idx is False."""
    _require_cuda(x)
    x, layout = as_layout(x)
    B, C_, H, W, _ = x.shape
    (kh, kw), (sh, sw), (ph, pw) = _pair(kernel), _pair(stride), _pair(padding)
    Ho, Wo = qmaxpool_out_hw(H, W, kernel, stride, padding)
    out = empty_q((B, C_, max(Ho, 1), max(Wo, 1), 4), x.dtype, x.device, layout)
    idx = empty_q((B, C_, max(Ho, 1), max(Wo, 1), 4), torch.uint8, x.device, layout) if with_idx else None
    check(_lib.load().quan_qmaxpool_fwd(x.data_ptr(), out.data_ptr(), idx.data_ptr() if with_idx else None, B, C_, H, W,
                                        kh, kw, sh, sw, ph, pw, _dtype_code(x), layout, _stream(x)), "quan_qmaxpool_fwd")
    return out, idx


def qmaxpool_bwd(dy: torch.Tensor, idx: torch.Tensor, in_hw, kernel, stride, padding) -> torch.Tensor:
    _require_cuda(dy, idx)
    layout = layout_of(idx)
    dy, _ = as_layout(dy, layout)
    B, C_, Ho, Wo, _ = dy.shape
    H, W = in_hw
    (kh, kw), (sh, sw), (ph, pw) = _pair(kernel), _pair(stride), _pair(padding)
    out = empty_q((B, C_, H, W, 4), dy.dtype, dy.device, layout)
    check(_lib.load().quan_qmaxpool_bwd(dy.data_ptr(), idx.data_ptr(), out.data_ptr(), B, C_, H, W, kh, kw, sh, sw, ph, pw,
                                        _dtype_code(dy), layout, _stream(dy)), "quan_qmaxpool_bwd")
    return out


# ---- QER (quaternion -> real extraction of the heads) -------------------------------------------------------------------------------
QER_MAX_K, QER_MAX_N = 256, 64


def qer_supported(C_: int, N: int, dtype: torch.dtype) -> bool:
    """Shapes quan_qer_* serve (include/quan_sm100.h): 4C <= 256 input and N <= 64 output channels; bf16 rows of whole 16-element
    MMA steps (C % 4 == 0) and, above 128 input channels, whole 128-column chunks."""
    K = 4 * C_
    if K > QER_MAX_K or N > QER_MAX_N or dtype not in (torch.float32, torch.bfloat16):
        return False
    return dtype == torch.float32 or (C_ % 4 == 0 and (K <= 128 or K % 128 == 0))


def qer_fwd(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
            col0: int = 0, writable: int = 0) -> torch.Tensor:
    """head.py:40-47 on a dense BHWQC activation [B,C,H,W,4].  Returns the logical [B,N,H,W] result in channels-last memory; with
    `out` (a [B,H,W,Ctot] contiguous buffer) the N columns starting at `col0` of every pixel row are written in place; `writable`
    (>= N) columns from col0 on belong to this call (row padding, zero-filled: lets a ragged width go out as whole vectors)."""
    _require_cuda(x, weight, bias)
    B, C_, H, W, _ = x.shape
    N = weight.size(0)
    w, b = _f32c(weight), _f32c(bias)
    buf = torch.empty((B, H, W, N), dtype=x.dtype, device=x.device) if out is None else out
    ld = buf.size(3)
    check(_lib.load().quan_qer_fwd(x.data_ptr(), w.data_ptr(), _ptr(b), buf.data_ptr() + col0 * buf.element_size(), B * H * W, C_, N, ld,
                                   int(writable), _dtype_code(x), _stream(x)), "quan_qer_fwd")
    return buf.permute(0, 3, 1, 2) if out is None else out


def pixel_rows(dy: torch.Tensor) -> torch.Tensor:
    """A logical [B,N,H,W] tensor whose memory is one row of N contiguous values per pixel, rows `stride(3)` elements apart:
    channels-last tensors, their channel slices (what CatBackward hands out) and padded-row tensors pass through; anything else is
    copied to channels-last."""
    B, N, H, W = dy.shape
    sb, sn, sh, sw = dy.stride()
    if (sn == 1 or N == 1) and sw >= N and sh == W * sw and sb == H * sh:
        return dy
    return dy.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


def qer_bwd(dy: torch.Tensor, x: torch.Tensor, weight: torch.Tensor, need_dx: bool, need_dw: bool, need_db: bool, readable: int = 0):
    """Returns (dx like x or None, dweight like weight (fp32) or None, dbias [N] or None); dy: logical [B,N,H,W] in x's dtype, pixel rows
    (see pixel_rows); `readable` (>= N) columns of every dy row may be read (finite row padding)."""
    _require_cuda(dy, x, weight)
    B, C_, H, W, _ = x.shape
    N = weight.size(0)
    dy = pixel_rows(dy)
    w = _f32c(weight)
    lib = _lib.load()
    dx = torch.empty_like(x, memory_format=torch.preserve_format) if need_dx else None
    dw = torch.empty_like(w) if (need_dw or need_db) else None
    db = torch.empty(N, dtype=torch.float32, device=x.device) if need_db else None
    npix = B * H * W
    wsb = _workspace(lib.quan_qer_workspace_bytes(npix, C_, N, _dtype_code(x)), x.device) if dw is not None else None
    if dy.stride(3) < max(readable, N):
        readable = 0
    side = _deferred["side"].get(x.device.index) if (_deferred["on"] and dw is not None) else None
    if side is None:
        check(lib.quan_qer_bwd(dy.data_ptr(), dy.stride(3), int(readable), x.data_ptr(), w.data_ptr(), _ptr(dx), _ptr(dw), _ptr(db), npix, C_, N,
                               _dtype_code(x), _ptr(wsb), 0 if wsb is None else wsb.numel(), _stream(x)), "quan_qer_bwd")
        return dx, (dw if need_dw else None), db
    # deferred_wgrad: dX on this stream, the weight / bias gradient (and its fold) on the lent side stream, joined with the others
    if dx is not None:
        check(lib.quan_qer_bwd(dy.data_ptr(), dy.stride(3), int(readable), x.data_ptr(), w.data_ptr(), dx.data_ptr(), None, None, npix, C_, N,
                               _dtype_code(x), None, 0, _stream(x)), "quan_qer_bwd")
    side.wait_stream(torch.cuda.current_stream(x.device))
    with torch.cuda.stream(side):
        wss = _workspace(lib.quan_qer_workspace_bytes(npix, C_, N, _dtype_code(x)), x.device)      # keyed by the (side) stream
        check(lib.quan_qer_bwd(dy.data_ptr(), dy.stride(3), int(readable), x.data_ptr(), w.data_ptr(), None, dw.data_ptr(), _ptr(db), npix, C_, N,
                               _dtype_code(x), wss.data_ptr(), wss.numel(), side.cuda_stream), "quan_qer_bwd")
    _keep_for_side(x.device, dy, x, dw, db)
    return dx, (dw if need_dw else None), db


# ---- QAttention core ------------------------------------------------------------------------------------------------
def qattention_fwd(qkv: torch.Tensor, heads: int, key_dim: int, head_dim: int, scale: float):
    """block.py:1520-1540 fused: qkv [B, heads*(2K+V), H, W, 4] (BHWQC) -> (o [B, heads*V, H, W, 4], lse [B*4*heads*H*W])."""
    _require_cuda(qkv)
    qkv, layout = as_layout(qkv, LAYOUT_BHWQC)
    B, Cq, H, W, _ = qkv.shape
    if Cq != heads * (2 * key_dim + head_dim):
        raise RuntimeError(f"qattention_fwd: {Cq} channels do not split into {heads} heads of q,k ({key_dim}) and v ({head_dim})")
    o = empty_q((B, heads * head_dim, H, W, 4), qkv.dtype, qkv.device, LAYOUT_BHWQC)
    lse = torch.empty(B * 4 * heads * H * W, dtype=torch.float32, device=qkv.device)
    check(_lib.load().quan_qattention_fwd(qkv.data_ptr(), o.data_ptr(), lse.data_ptr(), B, H, W, heads, key_dim, head_dim, float(scale),
                                          _dtype_code(qkv), layout, _stream(qkv)), "quan_qattention_fwd")
    return o, lse


def qattention_bwd(qkv: torch.Tensor, o: torch.Tensor, d_o: torch.Tensor, lse: torch.Tensor, heads: int, key_dim: int, head_dim: int,
                   scale: float) -> torch.Tensor:
    _require_cuda(qkv, o, d_o, lse)
    d_o, _ = as_layout(d_o, LAYOUT_BHWQC)
    if d_o.dtype != qkv.dtype:
        d_o = d_o.to(qkv.dtype)
    B, Cq, H, W, _ = qkv.shape
    dqkv = torch.empty_like(qkv, memory_format=torch.preserve_format)
    check(_lib.load().quan_qattention_bwd(qkv.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, H, W, heads,
                                          key_dim, head_dim, float(scale), _dtype_code(qkv), LAYOUT_BHWQC, _stream(qkv)),
          "quan_qattention_bwd")
    return dqkv


# ---- IQBN ----------------------------------------------------------------------------------------------------------
def iqbn_train_stats(x: torch.Tensor, layout: int, gamma: torch.Tensor, beta: torch.Tensor, eps: float, momentum: float,
                     running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor]) -> torch.Tensor:
    """One launch: per-(c,q) mean / var(+1e-8) / rstd + the scale/shift table into stats[20C]; updates running stats."""
    B, C_, H, W, _ = x.shape
    stats = torch.empty(20 * C_, dtype=torch.float32, device=x.device)
    ws = _iqbn_workspace(C_, x.device)
    check(_lib.load().quan_iqbn_train_stats(x.data_ptr(), B, C_, H, W, _dtype_code(x), layout, gamma.data_ptr(),
                                            beta.data_ptr(), eps, momentum, _ptr(running_mean), _ptr(running_var),
                                            stats.data_ptr(), ws.data_ptr(), ws.numel(), _stream(x)),
          "quan_iqbn_train_stats")
    return stats


def iqbn_partial_sums(x: torch.Tensor, layout: int) -> torch.Tensor:
    B, C_, H, W, _ = x.shape
    sums = torch.empty(8 * C_, dtype=torch.float64, device=x.device)
    ws = _iqbn_workspace(C_, x.device)
    check(_lib.load().quan_iqbn_partial_sums(x.data_ptr(), B, C_, H, W, _dtype_code(x), layout, sums.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _stream(x)), "quan_iqbn_partial_sums")
    return sums


def iqbn_finalize_stats(sums: torch.Tensor, count: float, C_: int, gamma, beta, eps: float, momentum: float,
                        running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor]) -> torch.Tensor:
    stats = torch.empty(20 * C_, dtype=torch.float32, device=sums.device)
    check(_lib.load().quan_iqbn_finalize_stats(sums.data_ptr(), float(count), C_, gamma.data_ptr(), beta.data_ptr(), eps,
                                               momentum, _ptr(running_mean), _ptr(running_var), stats.data_ptr(),
                                               _stream(sums)), "quan_iqbn_finalize_stats")
    return stats


def iqbn_eval_stats(gamma, beta, running_mean, running_var, eps: float) -> torch.Tensor:
    """The [20C] coefficient table of eval-mode IQBN (running statistics); feed it to iqbn_apply_fwd."""
    C_ = gamma.size(0)
    stats = torch.empty(20 * C_, dtype=torch.float32, device=gamma.device)
    check(_lib.load().quan_iqbn_eval_stats(gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr(),
                                           running_var.data_ptr(), eps, C_, stats.data_ptr(), _stream(gamma)),
          "quan_iqbn_eval_stats")
    return stats


def iqbn_apply_fwd(x: torch.Tensor, layout: int, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                   act: int) -> torch.Tensor:
    B, C_, H, W, _ = x.shape
    y = torch.empty_like(x, memory_format=torch.preserve_format)
    check(_lib.load().quan_iqbn_apply_fwd(x.data_ptr(), y.data_ptr(), B, C_, H, W, _dtype_code(x), layout,
                                          stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), act, _stream(x)),
          "quan_iqbn_apply_fwd")
    return y


def iqbn_eval_fwd(x: torch.Tensor, layout: int, gamma, beta, running_mean, running_var, eps: float,
                  act: int) -> torch.Tensor:
    B, C_, H, W, _ = x.shape
    y = torch.empty_like(x, memory_format=torch.preserve_format)
    check(_lib.load().quan_iqbn_eval_fwd(x.data_ptr(), y.data_ptr(), B, C_, H, W, _dtype_code(x), layout,
                                         gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr(),
                                         running_var.data_ptr(), eps, act, _stream(x)), "quan_iqbn_eval_fwd")
    return y


def iqbn_bwd_reduce(dy: torch.Tensor, x: torch.Tensor, layout: int, stats, gamma, beta, act: int,
                    count: float = 0.0) -> torch.Tensor:
    """sums[14C] doubles: {sum dz, sum dz*xhat} + (when count > 0) the backward coefficient table."""
    B, C_, H, W, _ = x.shape
    sums = torch.empty(14 * C_, dtype=torch.float64, device=x.device)
    ws = _iqbn_workspace(C_, x.device)
    check(_lib.load().quan_iqbn_bwd_reduce(dy.data_ptr(), x.data_ptr(), B, C_, H, W, _dtype_code(x), layout,
                                           stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), act, float(count),
                                           sums.data_ptr(), ws.data_ptr(), ws.numel(), _stream(x)), "quan_iqbn_bwd_reduce")
    return sums


def iqbn_bwd_coef(sums: torch.Tensor, count: float, stats, gamma) -> None:
    """Synced IQBN: build the backward coefficient table after sums[0:8C] were all-reduced."""
    C_ = gamma.size(0)
    check(_lib.load().quan_iqbn_bwd_coef(sums.data_ptr(), float(count), C_, stats.data_ptr(), gamma.data_ptr(),
                                         _stream(sums)), "quan_iqbn_bwd_coef")


def iqbn_bwd_apply(dy, x, layout: int, stats, gamma, beta, act: int, sums, count: float,
                   want_param_grads: bool = True, mix_t: Optional[Sequence[float]] = None):
    B, C_, H, W, _ = x.shape
    dx = torch.empty_like(x, memory_format=torch.preserve_format)
    dgamma = dbeta = None
    if want_param_grads:
        dgamma = torch.empty(C_, 4, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(C_, 4, dtype=torch.float32, device=x.device)
    mix_arg = None if mix_t is None else C.cast(_mix_arg(mix_t), C.c_void_p)
    check(_lib.load().quan_iqbn_bwd_apply(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), B, C_, H, W, _dtype_code(x),
                                          layout, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), act,
                                          sums.data_ptr(), float(count), _ptr(dgamma), _ptr(dbeta), mix_arg,
                                          _stream(x)), "quan_iqbn_bwd_apply")
    return dx, dgamma, dbeta


def iqbn_eval_bwd(dy, x, layout: int, gamma, beta, running_mean, running_var, eps: float, act: int) -> torch.Tensor:
    B, C_, H, W, _ = x.shape
    dx = torch.empty_like(x, memory_format=torch.preserve_format)
    check(_lib.load().quan_iqbn_eval_bwd(dy.data_ptr(), x.data_ptr(), dx.data_ptr(), B, C_, H, W, _dtype_code(x),
                                         layout, gamma.data_ptr(), beta.data_ptr(), running_mean.data_ptr(),
                                         running_var.data_ptr(), eps, act, _stream(x)), "quan_iqbn_eval_bwd")
    return dx


# ---- QConv2D ---------------------------------------------------------------------------------------------------------
_ws_bytes_cache = {}
_wants_mixed_cache = {}


def _conv_ws_bytes(lib, d: ConvDims, dtype_code: int, layout: int, algo: int) -> int:
    """quan_qconv2d_workspace_bytes, cached per shape (the C side plans all three passes to answer)."""
    key = (d.B, d.Ci, d.Co, d.H, d.W, d.kH, d.kW, d.sH, d.sW, d.pH, d.pW, d.dH, d.dW, d.groups, dtype_code, layout, algo)
    n = _ws_bytes_cache.get(key)
    if n is None:
        n = _ws_bytes_cache[key] = lib.quan_qconv2d_workspace_bytes(C.byref(d), dtype_code, layout, algo)
    return n


def conv_dims(x_shape, w_shape, stride, padding, dilation, groups) -> ConvDims:
    B, Ci, H, W, _ = x_shape
    Co, _, kH, kW = w_shape
    return ConvDims(B, Ci, Co, H, W, kH, kW, stride[0], stride[1], padding[0], padding[1], dilation[0], dilation[1],
                    groups)


def conv_out_shape(d: ConvDims) -> Tuple[int, int]:
    Ho = (d.H + 2 * d.pH - d.dH * (d.kH - 1) - 1) // d.sH + 1
    Wo = (d.W + 2 * d.pW - d.dW * (d.kW - 1) - 1) // d.sW + 1
    return Ho, Wo


def _weights_arg(ws: Sequence[torch.Tensor]):
    return PtrArray4(*[w.data_ptr() for w in ws])


def qconv2d_fwd(x: torch.Tensor, weights: Sequence[torch.Tensor], bias_r: Optional[torch.Tensor], stride, padding,
                dilation, groups: int, mix_matrix: Sequence[float], algo: int = ALGO_AUTO,
                layout: Optional[int] = None, with_stats: bool = False):
    """y = M (W_sep * x).  with_stats=True returns (y, nparts): the tensor-core epilogue also wrote per-CTA partial IQBN
    sums of y into the IQBN workspace of C_o (nparts slots; 0 = not produced, run iqbn_train_stats on y instead)."""
    _require_cuda(x, *weights, bias_r)
    x, layout = as_layout(x, layout)
    ws = [_f32c(w) for w in weights]
    bias_r = _f32c(bias_r)
    if ws[0].dim() != 4 or x.size(1) != ws[0].size(1) * groups:
        raise RuntimeError(f"qconv2d_fwd: input has {x.size(1)} quaternion channels, weight expects "
                           f"{ws[0].size(1) * groups} (weight {tuple(ws[0].shape)}, groups={groups})")
    d = conv_dims(x.shape, ws[0].shape, stride, padding, dilation, groups)
    Ho, Wo = conv_out_shape(d)
    if Ho <= 0 or Wo <= 0:
        raise RuntimeError(f"qconv2d_fwd: empty output for input {tuple(x.shape)} and kernel {tuple(ws[0].shape)}")
    y = empty_q((d.B, d.Co, Ho, Wo, 4), x.dtype, x.device, layout)
    lib = _lib.load()
    nws = _conv_ws_bytes(lib, d, _dtype_code(x), layout, algo)
    wsb = _workspace(nws, x.device)
    wa = _weights_arg(ws)
    if with_stats:
        iws = _iqbn_workspace(d.Co, x.device)
        nparts = C.c_int(0)
        check(lib.quan_qconv2d_fwd_stats(x.data_ptr(), C.cast(wa, C.c_void_p), _ptr(bias_r), y.data_ptr(), C.byref(d),
                                         _dtype_code(x), layout, C.cast(_mix_arg(mix_matrix), C.c_void_p), algo,
                                         wsb.data_ptr(), wsb.numel(), iws.data_ptr(), iws.numel(), C.byref(nparts),
                                         _stream(x)), "quan_qconv2d_fwd_stats")
        return y, nparts.value
    check(lib.quan_qconv2d_fwd(x.data_ptr(), C.cast(wa, C.c_void_p), _ptr(bias_r), y.data_ptr(), C.byref(d),
                               _dtype_code(x), layout, C.cast(_mix_arg(mix_matrix), C.c_void_p), algo, wsb.data_ptr(),
                               wsb.numel(), _stream(x)), "quan_qconv2d_fwd")
    return y


def iqbn_finalize_partials(nparts: int, count: float, C_: int, gamma, beta, eps: float, momentum: float,
                           running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor]) -> torch.Tensor:
    """Second half of iqbn_train_stats on partial sums a conv epilogue left in the IQBN workspace (qconv2d_fwd with_stats)."""
    stats = torch.empty(20 * C_, dtype=torch.float32, device=gamma.device)
    ws = _iqbn_workspace(C_, gamma.device)
    check(_lib.load().quan_iqbn_finalize_partials(ws.data_ptr(), int(nparts), float(count), C_, gamma.data_ptr(),
                                                  beta.data_ptr(), eps, momentum, _ptr(running_mean), _ptr(running_var),
                                                  stats.data_ptr(), _stream(gamma)), "quan_iqbn_finalize_partials")
    return stats


def qconv2d_bwd(dy: torch.Tensor, x: torch.Tensor, weights: Sequence[torch.Tensor], stride, padding, dilation,
                groups: int, mix_matrix: Sequence[float], need_dx: bool = True, need_dw: bool = True,
                need_db: bool = False, algo: int = ALGO_AUTO, premixed: bool = False):
    """Returns (dx or None, [dw_r, dw_i, dw_j, dw_k] or None, db_r or None).  `premixed`: dy already holds G = M^T dY
    (emitted by iqbn_bwd_apply(mix_t=...)); legal only when qconv2d_bwd_wants_mixed() says so."""
    _require_cuda(dy, x, *weights)
    # the gradient decides the layout: a C_q = 1 input (first layer) is the same bytes in both layouts
    dy, layout = as_layout(dy)
    x, _ = as_layout(x, layout)
    ws = [_f32c(w) for w in weights]
    d = conv_dims(x.shape, ws[0].shape, stride, padding, dilation, groups)
    lib = _lib.load()
    dx = torch.empty_like(x, memory_format=torch.preserve_format) if need_dx else None
    dws = [torch.empty_like(w) for w in ws] if need_dw else None
    db = torch.empty(d.Co, dtype=torch.float32, device=x.device) if need_db else None
    nws = _conv_ws_bytes(lib, d, _dtype_code(x), layout, algo)
    wsb = _workspace(nws, x.device)
    wa = _weights_arg(ws)
    dwa = None if dws is None else _weights_arg(dws)
    fn = lib.quan_qconv2d_bwd_premixed if premixed else lib.quan_qconv2d_bwd
    check(fn(dy.data_ptr(), x.data_ptr(), C.cast(wa, C.c_void_p), _ptr(dx),
             None if dwa is None else C.cast(dwa, C.c_void_p), _ptr(db), C.byref(d), _dtype_code(x),
             layout, C.cast(_mix_arg(mix_matrix), C.c_void_p), algo, wsb.data_ptr(), wsb.numel(),
             _stream(x)), "quan_qconv2d_bwd")
    if dws is not None:
        _keep_for_side(x.device, dy, x, *dws)
    return dx, dws, db


def qconv2d_bwd_wants_mixed(x_shape, w_shape, stride, padding, dilation, groups, dtype: torch.dtype, layout: int,
                            algo: int = ALGO_AUTO, need_dx: bool = True, need_dw: bool = True) -> bool:
    """True when every requested backward pass of this conv reads G = M^T dY (so the IQBN backward can emit G directly)."""
    key = (tuple(x_shape), tuple(w_shape), tuple(stride), tuple(padding), tuple(dilation), groups, dtype, layout, algo,
           bool(need_dx), bool(need_dw))
    r = _wants_mixed_cache.get(key)
    if r is None:
        d = conv_dims(x_shape, w_shape, stride, padding, dilation, groups)
        r = _wants_mixed_cache[key] = _lib.load().quan_qconv2d_bwd_wants_mixed(
            C.byref(d), F32 if dtype == torch.float32 else BF16, layout, algo, int(need_dx), int(need_dw)) == 1
    return r


def qconv2d_pick_algo(x_shape, w_shape, stride, padding, dilation, groups, dtype: torch.dtype, layout: int,
                      pass_: int = 0) -> int:
    d = conv_dims(x_shape, w_shape, stride, padding, dilation, groups)
    return _lib.load().quan_qconv2d_pick_algo(C.byref(d), F32 if dtype == torch.float32 else BF16, layout, pass_)


# ---- Conv block (conv -> IQBN(batch stats) -> act) in one C call per direction ------------------------------------------
def conv_block_fwd(x: torch.Tensor, weights: Sequence[torch.Tensor], gamma, beta, running_mean, running_var, stride, padding,
                   dilation, groups: int, mix_matrix: Sequence[float], algo: int, eps: float, momentum: float, act: int,
                   layout: int, epilogue_stats: bool = True):
    """Returns (y = conv output, out = act(IQBN(y)), stats[20*C_o]); x must already be dense in `layout`."""
    ws = [_f32c(w) for w in weights]
    d = conv_dims(x.shape, ws[0].shape, stride, padding, dilation, groups)
    Ho, Wo = conv_out_shape(d)
    if Ho <= 0 or Wo <= 0 or x.size(1) != ws[0].size(1) * groups:
        raise RuntimeError(f"conv_block_fwd: input {tuple(x.shape)} does not fit weight {tuple(ws[0].shape)} (groups={groups})")
    y = empty_q((d.B, d.Co, Ho, Wo, 4), x.dtype, x.device, layout)
    out = torch.empty_like(y, memory_format=torch.preserve_format)
    stats = torch.empty(20 * d.Co, dtype=torch.float32, device=x.device)
    lib = _lib.load()
    code = _dtype_code(x)
    wsb = _workspace(_conv_ws_bytes(lib, d, code, layout, algo), x.device)
    iws = _iqbn_workspace(d.Co, x.device)
    wa = _weights_arg(ws)
    check(lib.quan_conv_block_fwd(x.data_ptr(), C.cast(wa, C.c_void_p), gamma.data_ptr(), beta.data_ptr(), _ptr(running_mean),
                                  _ptr(running_var), y.data_ptr(), out.data_ptr(), stats.data_ptr(), C.byref(d), code, layout,
                                  C.cast(_mix_arg(mix_matrix), C.c_void_p), algo, eps, momentum, act, int(epilogue_stats),
                                  wsb.data_ptr(), wsb.numel(), iws.data_ptr(), iws.numel(), _stream(x)), "quan_conv_block_fwd")
    return y, out, stats


def conv_block_bwd(dout: torch.Tensor, x: torch.Tensor, y: torch.Tensor, weights: Sequence[torch.Tensor], stats, gamma, beta,
                   stride, padding, dilation, groups: int, mix_matrix: Sequence[float], algo: int, act: int, layout: int,
                   need_dx: bool, need_dw: bool):
    """Returns (dx or None, [dw x4] or None, dgamma, dbeta); dout / x / y dense in `layout`."""
    ws = [_f32c(w) for w in weights]
    d = conv_dims(x.shape, ws[0].shape, stride, padding, dilation, groups)
    lib = _lib.load()
    code = _dtype_code(x)
    g = torch.empty_like(y, memory_format=torch.preserve_format)
    dx = torch.empty_like(x, memory_format=torch.preserve_format) if need_dx else None
    dws = [torch.empty_like(w) for w in ws] if need_dw else None
    dgamma = torch.empty(d.Co, 4, dtype=torch.float32, device=x.device)
    dbeta = torch.empty(d.Co, 4, dtype=torch.float32, device=x.device)
    sums = torch.empty(14 * d.Co, dtype=torch.float64, device=x.device)
    wsb = _workspace(_conv_ws_bytes(lib, d, code, layout, algo), x.device)
    iws = _iqbn_workspace(d.Co, x.device)
    wa = _weights_arg(ws)
    dwa = None if dws is None else _weights_arg(dws)
    check(lib.quan_conv_block_bwd(dout.data_ptr(), x.data_ptr(), y.data_ptr(), C.cast(wa, C.c_void_p), stats.data_ptr(),
                                  gamma.data_ptr(), beta.data_ptr(), g.data_ptr(), _ptr(dx),
                                  None if dwa is None else C.cast(dwa, C.c_void_p), dgamma.data_ptr(), dbeta.data_ptr(),
                                  sums.data_ptr(), C.byref(d), code, layout, C.cast(_mix_arg(mix_matrix), C.c_void_p), algo, act,
                                  wsb.data_ptr(), wsb.numel(), iws.data_ptr(), iws.numel(), _stream(x)), "quan_conv_block_bwd")
    if dws is not None:
        _keep_for_side(x.device, g, x, *dws)
    return dx, dws, dgamma, dbeta


def conv_block_eval_fwd(x: torch.Tensor, weights: Sequence[torch.Tensor], gamma, beta, running_mean, running_var, stride,
                        padding, dilation, groups: int, mix_matrix: Sequence[float], algo: int, eps: float, act: int,
                        layout: int) -> torch.Tensor:
    """Inference `Conv` block: act(IQBN_eval(QConv2D(x))); x must already be dense in `layout`."""
    ws = [_f32c(w) for w in weights]
    d = conv_dims(x.shape, ws[0].shape, stride, padding, dilation, groups)
    Ho, Wo = conv_out_shape(d)
    if Ho <= 0 or Wo <= 0 or x.size(1) != ws[0].size(1) * groups:
        raise RuntimeError(f"conv_block_eval_fwd: input {tuple(x.shape)} does not fit weight {tuple(ws[0].shape)} (groups={groups})")
    lib = _lib.load()
    code = _dtype_code(x)
    out = empty_q((d.B, d.Co, Ho, Wo, 4), x.dtype, x.device, layout)
    stats = torch.empty(20 * d.Co, dtype=torch.float32, device=x.device)
    fused = algo in (ALGO_AUTO, ALGO_TCGEN05) and lib.quan_qconv2d_pick_algo(C.byref(d), code, layout, 0) == ALGO_TCGEN05
    y = None if fused else torch.empty_like(out, memory_format=torch.preserve_format)
    wsb = _workspace(_conv_ws_bytes(lib, d, code, layout, algo), x.device)
    wa = _weights_arg(ws)
    check(lib.quan_conv_block_eval_fwd(x.data_ptr(), C.cast(wa, C.c_void_p), gamma.data_ptr(), beta.data_ptr(),
                                       running_mean.data_ptr(), running_var.data_ptr(), out.data_ptr(), stats.data_ptr(),
                                       _ptr(y), C.byref(d), code, layout, C.cast(_mix_arg(mix_matrix), C.c_void_p), algo, eps,
                                       act, wsb.data_ptr(), wsb.numel(), _stream(x)), "quan_conv_block_eval_fwd")
    return out


# ---- deferred weight-gradient join (include/quan_sm100.h: quan_bwd_side_stream_set / _join) --------------------------------------------------
_deferred = {"on": False, "side": {}, "ws": {}}
DEFERRED_WGRAD_WS_BYTES = 256 << 20


class deferred_wgrad:
    """Context manager around a backward pass: the tensor-core wgrad chains of the narrow layers are forked onto one lent side stream and
    joined ONCE, at exit (or at `ops.deferred_wgrad_join()`): they only feed the optimizer, so they leave the dX critical path.  Tensors a
    forked wgrad touches are handed to the allocator with record_stream.  QUAN_BWD_DEFER=0 turns it into a no-op."""

    def __init__(self, device: torch.device):
        import os
        self.device = torch.device(device)
        self.enabled = os.environ.get("QUAN_BWD_DEFER", "1") != "0" and self.device.type == "cuda"

    def __enter__(self):
        if not self.enabled:
            return self
        key = self.device.index if self.device.index is not None else torch.cuda.current_device()
        if key not in _deferred["side"]:
            _deferred["side"][key] = torch.cuda.Stream(device=self.device)
            _deferred["ws"][key] = torch.empty(DEFERRED_WGRAD_WS_BYTES, dtype=torch.uint8, device=self.device)
        side, ws = _deferred["side"][key], _deferred["ws"][key]
        with torch.cuda.device(self.device):
            check(_lib.load().quan_bwd_side_stream_set(side.cuda_stream, ws.data_ptr(), ws.numel()), "quan_bwd_side_stream_set")
        _deferred["on"] = True
        return self

    def __exit__(self, *exc):
        if not self.enabled:
            return False
        deferred_wgrad_join(self.device)
        _deferred["on"] = False
        with torch.cuda.device(self.device):
            _lib.load().quan_bwd_side_stream_set(None, None, 0)
        return False


def deferred_wgrad_join(device=None) -> None:
    """The current stream waits for every wgrad forked so far (no-op outside `deferred_wgrad`)."""
    if not _deferred["on"]:
        return
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        check(_lib.load().quan_bwd_side_stream_join(torch.cuda.current_stream(dev).cuda_stream), "quan_bwd_side_stream_join")


def _keep_for_side(device: torch.device, *tensors) -> None:
    """A forked wgrad may still read / write these when the caller drops them: the allocator must not hand their memory out before the
    side stream has passed this point."""
    if not _deferred["on"]:
        return
    side = _deferred["side"].get(device.index if device.index is not None else torch.cuda.current_device())
    if side is None:
        return
    for t in tensors:
        if t is not None:
            t.record_stream(side)


# ---- pack plan (include/quan_sm100.h): every packed weight of a step rebuilt by ONE launch ------------------------------------------------
class PackPlan:
    """with-less protocol used by graphs.GraphedTrainStep:  p = PackPlan(); p.record(); <one eager step>; p.commit(device);
    then `p.run()` as the first thing of every step (captured: first node of the forward graph); p.release() when the convolutions
    should pack for themselves again.  Holds the arena and the job table (they must outlive every graph captured under the plan)."""

    def __init__(self):
        self.arena = self.table = None
        self.jobs = 0

    def record(self) -> None:
        check(min(0, _lib.load().quan_pack_plan_record(1)), "quan_pack_plan_record")

    def commit(self, device: torch.device) -> int:
        lib = _lib.load()
        self.jobs = lib.quan_pack_plan_record(0)
        if self.jobs <= 0:
            return 0
        tb = C.c_size_t(0)
        nbytes = lib.quan_pack_plan_bytes(C.byref(tb))
        self.arena = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        self.table = torch.empty(tb.value, dtype=torch.uint8, device=device)
        base = (self.arena.data_ptr() + 1023) // 1024 * 1024
        check(lib.quan_pack_plan_commit(base, nbytes, self.table.data_ptr(), tb.value, torch.cuda.current_stream(device).cuda_stream),
              "quan_pack_plan_commit")
        return self.jobs

    def run(self) -> None:
        if self.jobs > 0:
            check(_lib.load().quan_pack_plan_run(torch.cuda.current_stream(self.arena.device).cuda_stream), "quan_pack_plan_run")

    def release(self) -> None:
        _lib.load().quan_pack_plan_release()


# ---- device activation -------------------------------------------------------------------------------------------------------------
# The library launches on the calling thread's CURRENT device (raw pointers + a stream handle carry no device); a tensor on cuda:1
# handed over while cuda:0 is current would meet the wrong stream table.  Every entry point that launches therefore runs under the
# device of its first CUDA tensor argument; when that already is the current device (one process per GPU — the normal case) the guard
# costs one integer comparison.  (C side: function attributes / occupancy answers are latched per device, common.cuh DeviceOnce.)
def _device_guarded(fn):
    import functools

    @functools.wraps(fn)
    def run(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor):
                if a.is_cuda and a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)

    return run


for _name in ("poincare_fwd", "poincare_bwd", "convert_layout", "mix", "qupsample_fwd", "qupsample_bwd", "qmaxpool_fwd", "qmaxpool_bwd",
              "qattention_fwd", "qattention_bwd", "qer_fwd", "qer_bwd", "iqbn_train_stats", "iqbn_partial_sums", "iqbn_finalize_stats", "iqbn_eval_stats",
              "iqbn_apply_fwd", "iqbn_eval_fwd", "iqbn_bwd_reduce", "iqbn_bwd_coef", "iqbn_bwd_apply", "iqbn_eval_bwd", "qconv2d_fwd",
              "iqbn_finalize_partials", "qconv2d_bwd", "conv_block_fwd", "conv_block_bwd", "conv_block_eval_fwd", "as_layout"):
    globals()[_name] = _device_guarded(globals()[_name])
del _name
