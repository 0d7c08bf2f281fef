"""Build libquan_sm100.so in-tree with nvcc for sm_100a (no torch dependency in the library itself).

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  `python -m quan_ultralytics_b200.build`
rebuilds when any source is newer than the library.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIB = PKG / "libquan_sm100.so"
OBJ = PKG / "build"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-I", str(INCLUDE),
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    files = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h"))
    return max(f.stat().st_mtime for f in files)


def needs_build() -> bool:
    return (not LIB.exists()) or LIB.stat().st_mtime < _deps_mtime()


def _compile(src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    hdr_m = max(f.stat().st_mtime for f in list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h")))
    if obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_m):
        return obj
    cmd = [NVCC, *FLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose and r.stderr:
        print(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB
    OBJ.mkdir(exist_ok=True)
    if force:
        for o in OBJ.glob("*.o"):
            o.unlink()
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), sources()))
    cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart_static", "-ldl", "-lpthread", "-lrt",
           "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
