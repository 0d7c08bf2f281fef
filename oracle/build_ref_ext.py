"""Recipe for the reference's own CUDA extension (test / baseline infrastructure only — never on the product path).

    python oracle/build_ref_ext.py        # build container only: compiles the sources WHERE THEY LIE under /root/reference

Sources: /root/reference/ultralytics/nn/cuda/quaternion_ops_py.cpp + quaternion_ops.cu (the module the reference's
conv.py:47-60 imports as `quaternion_ops`; no build script ships with the reference).  Output: oracle/_ref/quaternion_ops.so
(git-ignored, travels to the GPU box with the snapshot).  Compiled for compute_100 (B200) with torch's own cpp_extension
machinery; nothing is copied into the repository.  bench.py times it as the `reference_cuda_ext` baseline
(SURVEY §8(d)); it computes mixing matrix M_B in fp32.
"""
from __future__ import annotations

import os
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/ultralytics/nn/cuda")
OUT = ROOT / "oracle" / "_ref"


def build(verbose: bool = False) -> Path | None:
    if not REF.exists():
        return None
    target = OUT / "quaternion_ops.so"
    if target.exists():
        return target
    OUT.mkdir(parents=True, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils import cpp_extension
    tmp = OUT / "_build"
    tmp.mkdir(exist_ok=True)
    cpp_extension.load(name="quaternion_ops", sources=[str(REF / "quaternion_ops_py.cpp"), str(REF / "quaternion_ops.cu")],
                       build_directory=str(tmp), extra_cuda_cflags=["-O3"], verbose=verbose, is_python_module=False)
    built = tmp / "quaternion_ops.so"
    shutil.copy2(built, target)
    shutil.rmtree(tmp, ignore_errors=True)
    return target


def load():
    """Import the built extension as a Python module (GPU box: only the prebuilt .so is used)."""
    target = OUT / "quaternion_ops.so"
    if not target.exists():
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location("quaternion_ops", str(target))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(verbose="-v" in sys.argv)
    print(p if p else "reference checkout not present: nothing built")
