"""The C-ABI library loads on a CPU-only host and exports every symbol include/quan_sm100.h declares."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "quan_sm100.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    names = re.findall(r"^\s*(?:int|size_t|uint64_t|const char\*)\s+(quan_[a-z0-9_]+)\s*\(", text, flags=re.M)
    assert len(names) >= 20
    return sorted(set(names))


def test_library_builds_and_loads():
    from quan_ultralytics_b200 import _lib
    lib = _lib.load()
    assert lib.quan_version() == 1
    assert b"sm_100a" in lib.quan_build_info()


@pytest.mark.parametrize("name", declared_functions())
def test_symbol_exported_and_bound(name):
    from quan_ultralytics_b200 import _lib
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    assert hasattr(raw, name), f"{name} declared in the header but not exported by the .so"
    assert name in _lib.PROTOTYPES, f"{name} has no ctypes prototype in _lib.py"


def test_no_undeclared_prototypes():
    from quan_ultralytics_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_functions()


def test_argument_errors_use_the_c_convention():
    """Host-side validation runs before any launch, so it is testable without a GPU."""
    from quan_ultralytics_b200 import _lib
    lib = _lib.load()
    d = _lib.ConvDims(1, 4, 6, 8, 8, 3, 3, 1, 1, 1, 1, 1, 1, 4)   # Co=6 not divisible by groups=4
    rc = lib.quan_qconv2d_fwd(None, None, None, None, ctypes.byref(d), 0, 0, None, 0, None, 0, None)
    assert rc == -1 and b"null" in lib.quan_last_error()
    mix = (ctypes.c_float * 16)(*([1.0] * 16))
    rc = lib.quan_qconv2d_fwd(1, 1, None, 1, ctypes.byref(d), 0, 0, ctypes.cast(mix, ctypes.c_void_p), 0, None, 0, None)
    assert rc == -2 and b"groups" in lib.quan_last_error()
    assert lib.quan_iqbn_train_stats(None, 1, 1, 1, 1, 0, 0, None, None, 1e-5, 0.1, None, None, 1, None, 0, None) == -1
    assert lib.quan_qupsample_nearest_fwd(1, 1, 1, 1, 1, 1, 0, 0, 0, None) == -1
    with pytest.raises(RuntimeError, match="argument/shape error"):
        _lib.check(-2, "quan_qconv2d_fwd")
    assert lib.quan_iqbn_workspace_bytes(16) == (592 * 8 * 16 + 8 * 16 + 2) * 8      # 4 x 148 row-split partials of [8C] fp64 + the single-launch accumulators and ticket
    assert lib.quan_qconv2d_pick_algo(ctypes.byref(_lib.ConvDims(1, 4, 4, 8, 8, 3, 3, 1, 1, 1, 1, 1, 1, 1)), 1, 0, 0) == 1


def test_engine_selection_is_host_logic():
    """quan_qconv2d_pick_algo needs no GPU: the small-channel engine takes the stems, including the Q-ResNet-34 7x7 stem
    (1 -> 16) in chunks of 8 output channels for forward and wgrad; its dgrad (never needed: the input is the image) and
    everything the special engines refuse stay on the generic CUDA-core engine."""
    from quan_ultralytics_b200 import _lib
    lib = _lib.load()
    pick = lambda dims, ps: lib.quan_qconv2d_pick_algo(ctypes.byref(_lib.ConvDims(*dims)), 1, 1, ps)   # bf16, BHWQC
    stem34 = (256, 1, 16, 224, 224, 7, 7, 2, 2, 3, 3, 1, 1, 1)
    assert [pick(stem34, ps) for ps in range(3)] == [4, 1, 4]
    yolo_stem = (16, 1, 4, 1024, 1024, 3, 3, 2, 2, 1, 1, 1, 1, 1)
    assert [pick(yolo_stem, ps) for ps in range(3)] == [4, 4, 4]
    dw = (16, 32, 32, 64, 64, 3, 3, 1, 1, 1, 1, 1, 1, 32)
    assert [pick(dw, ps) for ps in range(3)] == [3, 3, 3]
    odd = (2, 3, 5, 16, 16, 3, 3, 1, 1, 1, 1, 1, 1, 1)
    assert [pick(odd, ps) for ps in range(3)] == [1, 1, 1]

