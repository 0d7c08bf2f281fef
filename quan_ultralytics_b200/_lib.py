"""ctypes binding of libquan_sm100.so (the C ABI declared in include/quan_sm100.h).

There is deliberately NO fallback: if the shared library is missing or fails to load, importing the ops raises.
(BASELINE.json north_star: "no Triton, no multi-backend dispatch and no CPU fallback".)
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libquan_sm100.so"

F32, BF16 = 0, 1
LAYOUT_BCHWQ, LAYOUT_BHWQC = 0, 1
ACT_NONE, ACT_SILU = 0, 1
ALGO_AUTO, ALGO_DIRECT, ALGO_TCGEN05, ALGO_DEPTHWISE, ALGO_SMALLC = 0, 1, 2, 3, 4


class ConvDims(C.Structure):
    """struct quan_conv_dims (include/quan_sm100.h)."""

    _fields_ = [(n, C.c_int32) for n in
                ("B", "Ci", "Co", "H", "W", "kH", "kW", "sH", "sW", "pH", "pW", "dH", "dW", "groups")]


_vp, _i32, _int, _f, _d, _sz = C.c_void_p, C.c_int32, C.c_int, C.c_float, C.c_double, C.c_size_t
_pdims = C.POINTER(ConvDims)
PtrArray4 = C.c_void_p * 4  # `const float* const w[4]` / `float* const dw[4]`

# name -> (restype, argtypes); must list EVERY symbol declared in include/quan_sm100.h (tests check this).
PROTOTYPES = {
    "quan_version": (_int, []),
    "quan_last_error": (C.c_char_p, []),
    "quan_build_info": (C.c_char_p, []),
    "quan_launch_count": (C.c_uint64, []),
    "quan_kernel_timing_enable": (_int, [_int]),
    "quan_kernel_timing_report": (_sz, [C.c_char_p, _sz]),
    "quan_poincare_fwd": (_int, [_vp, _vp, _i32, _i32, _i32, _int, _vp]),
    "quan_poincare_bwd": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _int, _vp]),
    "quan_iqbn_workspace_bytes": (_sz, [_i32]),
    "quan_iqbn_train_stats": (_int, [_vp, _i32, _i32, _i32, _i32, _int, _int, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp, _sz, _vp]),
    "quan_iqbn_partial_sums": (_int, [_vp, _i32, _i32, _i32, _i32, _int, _int, _vp, _vp, _sz, _vp]),
    "quan_iqbn_finalize_stats": (_int, [_vp, _d, _i32, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp]),
    "quan_iqbn_eval_stats": (_int, [_vp, _vp, _vp, _vp, _f, _i32, _vp, _vp]),
    "quan_iqbn_bwd_coef": (_int, [_vp, _d, _i32, _vp, _vp, _vp]),
    "quan_iqbn_apply_fwd": (_int, [_vp, _vp, _i32, _i32, _i32, _i32, _int, _int, _vp, _vp, _vp, _int, _vp]),
    "quan_iqbn_eval_fwd": (_int, [_vp, _vp, _i32, _i32, _i32, _i32, _int, _int, _vp, _vp, _vp, _vp, _f, _int, _vp]),
    "quan_iqbn_bwd_reduce": (_int, [_vp, _vp, _i32, _i32, _i32, _i32, _int, _int, _vp, _vp, _vp, _int, _d, _vp, _vp, _sz, _vp]),
    "quan_iqbn_bwd_apply": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _int, _int, _vp, _vp, _vp, _int, _vp, _d,
                                   _vp, _vp, _vp, _vp]),
    "quan_iqbn_eval_bwd": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _int, _int, _vp, _vp, _vp, _vp, _f, _int, _vp]),
    "quan_qupsample_nearest_fwd": (_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _int, _int, _vp]),
    "quan_qupsample_nearest_bwd": (_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _int, _int, _vp]),
    "quan_qmaxpool_fwd": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _int, _int, _vp]),
    "quan_qmaxpool_bwd": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _int, _int, _vp]),
    "quan_mix": (_int, [_vp, _vp, _i32, _i32, _i32, _i32, _int, _int, _vp, _vp]),
    "quan_layout_convert": (_int, [_vp, _int, _vp, _int, _i32, _i32, _i32, _i32, _int, _vp]),
    "quan_qconv2d_workspace_bytes": (_sz, [_pdims, _int, _int, _int]),
    "quan_qconv2d_fwd": (_int, [_vp, _vp, _vp, _vp, _pdims, _int, _int, _vp, _int, _vp, _sz, _vp]),
    "quan_qconv2d_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _pdims, _int, _int, _vp, _int, _vp, _sz, _vp]),
    "quan_qconv2d_pick_algo": (_int, [_pdims, _int, _int, _int]),
    "quan_qconv2d_fwd_stats": (_int, [_vp, _vp, _vp, _vp, _pdims, _int, _int, _vp, _int, _vp, _sz, _vp, _sz, C.POINTER(C.c_int), _vp]),
    "quan_iqbn_finalize_partials": (_int, [_vp, _i32, _d, _i32, _vp, _vp, _f, _f, _vp, _vp, _vp, _vp]),
    "quan_conv_block_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _pdims, _int, _int, _vp, _int, _f, _f, _int, _int,
                                   _vp, _sz, _vp, _sz, _vp]),
    "quan_conv_block_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _pdims, _int, _int, _vp, _int,
                                   _int, _vp, _sz, _vp, _sz, _vp]),
    "quan_conv_block_eval_fwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _pdims, _int, _int, _vp, _int, _f, _int,
                                        _vp, _sz, _vp]),
    "quan_qconv2d_bwd_premixed": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _pdims, _int, _int, _vp, _int, _vp, _sz, _vp]),
    "quan_qconv2d_bwd_wants_mixed": (_int, [_pdims, _int, _int, _int, _int, _int]),
    "quan_bwd_side_stream_set": (_int, [_vp, _vp, _sz]),
    "quan_bwd_side_stream_join": (_int, [_vp]),
    "quan_pack_plan_record": (_int, [_int]),
    "quan_pack_plan_bytes": (_sz, [C.POINTER(C.c_size_t)]),
    "quan_pack_plan_commit": (_int, [_vp, _sz, _vp, _sz, _vp]),
    "quan_pack_plan_run": (_int, [_vp]),
    "quan_pack_plan_release": (_int, []),
    "quan_rows_cat": (_int, [_vp, _i32, _vp, C.c_int64, C.c_int64, _vp]),
    "quan_rows_split": (_int, [_vp, _i32, _vp, C.c_int64, C.c_int64, _vp]),
    "quan_rows_gather": (_int, [_vp, _vp, C.c_int64, _i32, C.c_int64, _vp]),
    "quan_qer_workspace_bytes": (_sz, [C.c_int64, _i32, _i32, _int]),
    "quan_qer_fwd": (_int, [_vp, _vp, _vp, _vp, C.c_int64, _i32, _i32, C.c_int64, _i32, _int, _vp]),
    "quan_qer_bwd": (_int, [_vp, C.c_int64, _i32, _vp, _vp, _vp, _vp, _vp, C.c_int64, _i32, _i32, _int, _vp, _sz, _vp]),
    "quan_qattention_fwd": (_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f, _int, _int, _vp]),
    "quan_qattention_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _f, _int, _int, _vp]),
    "quan_rotated_tal_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "quan_rotated_tal_assign": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _f, _f, _f, _vp, _vp, _vp, _vp,
                                       _vp, _sz, _vp]),
    "quan_obb_decode": (_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _int, _vp]),
    "quan_obb_loss_fwd_bwd": (_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _f, _f, _f, _f, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp]),
    "quan_sgd_clip_step": (_int, [_vp, _int, _vp, _vp, _int, _vp, _vp, _int, _vp]),
    "quan_ema_update": (_int, [_vp, _int, _vp, _vp, _vp]),
}

class CatSrc(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld_bytes", C.c_int64), ("row_bytes", C.c_int64)]


CAT_MAX = 8

_lib = None


class QuanLibraryError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the library (once).  Raises QuanLibraryError when it is absent — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if os.environ.get("QUAN_AUTO_BUILD", "1") == "1":
            from . import build as _build
            _build.build()
        if not LIB_PATH.exists():
            raise QuanLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m quan_ultralytics_b200.build` "
                "(there is no CPU or PyTorch fallback for the quaternion ops)")
    try:
        lib = C.CDLL(str(LIB_PATH))
    except OSError as e:  # pragma: no cover
        raise QuanLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise QuanLibraryError(f"{LIB_PATH} does not export {name}; rebuild the library") from e
        fn.restype = res
        fn.argtypes = args
    if lib.quan_version() != 1:
        raise QuanLibraryError(f"ABI mismatch: library reports version {lib.quan_version()}, binding expects 1")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    """Map the C return convention onto RuntimeError (the reference raises RuntimeError via TORCH_CHECK)."""
    if rc != 0:
        msg = load().quan_last_error().decode("utf-8", "replace")
        kind = "argument/shape error" if rc < 0 else "CUDA error"
        raise RuntimeError(f"{what}: {kind} {rc}: {msg}")
