"""Recipe: install the UNMODIFIED reference (bryceag11/QUAN_ultralytics) into git-ignored baseline/_ref/ so it travels to
the GPU box with the gpurun snapshot (like oracle/_ref/quaternion_ops.so does).

    python baseline/install_ref.py            # build container only (needs /root/reference); idempotent

What lands in baseline/_ref/ (nothing of it is tracked by git, nothing is edited):
  * `ultralytics/`            — `pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of the tree>`
                                 (the build writes egg-info next to pyproject.toml, /root/reference is read-only, hence the /tmp copy;
                                 --no-deps: matplotlib / seaborn / ultralytics-thop are not in the wheelhouse — quan_ultralytics_b200/refenv.py
                                 stubs the two that sit on the import path)
  * `classification/{quaternion,models}/` — the classification half is a script directory, not a package (no pyproject): its two
                                 library directories are installed by file copy; Q-WRN-16-2 / Q-ResNet-34 live there

Users of the install: tests/test_gpu_models.py (the real model graphs on the B200, swapped vs unswapped), bench.py (the
QUAN-YOLO11n-OBB training step: `--impl ours` = the reference's graph with the B200 modules installed, `--impl reference` = the same
graph untouched on the host cores).
"""
from __future__ import annotations

import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference")
OUT = ROOT / "baseline" / "_ref"


def installed() -> bool:
    return (OUT / "ultralytics" / "nn" / "modules" / "conv.py").exists() and (OUT / "classification" / "quaternion" / "qconv.py").exists()


def install(force: bool = False) -> Path | None:
    if not SRC.exists():
        return OUT if installed() else None
    if installed() and not force:
        return OUT
    if OUT.exists():
        shutil.rmtree(OUT)
    OUT.mkdir(parents=True)
    with tempfile.TemporaryDirectory(prefix="quan_ref_") as tmp:
        tmp = Path(tmp)
        shutil.copytree(SRC / "ultralytics", tmp / "ultralytics", ignore=shutil.ignore_patterns("__pycache__"))
        for f in ("pyproject.toml", "README.md", "LICENSE"):
            shutil.copy2(SRC / f, tmp / f)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", str(OUT), str(tmp)]
        subprocess.run(cmd, check=True)
    for sub in ("quaternion", "models"):
        shutil.copytree(SRC / "classification" / sub, OUT / "classification" / sub, ignore=shutil.ignore_patterns("__pycache__", "cuda"))
    shutil.rmtree(OUT / "bin", ignore_errors=True)
    assert installed()
    return OUT


if __name__ == "__main__":
    p = install(force="--force" in sys.argv)
    print(p if p else "no reference checkout here and nothing installed")
