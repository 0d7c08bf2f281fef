"""Locate and import a checkout / install of the reference (bryceag11/QUAN_ultralytics) — host plumbing for the drop-in.

The product is a plug-in for the reference: its model graphs (yolo11*-obb-quan.yaml via `ultralytics.nn.tasks`, Q-WRN / Q-ResNet via
`classification/models`) are the reference's own Python and are never rebuilt here.  This module finds the
reference tree, puts it on `sys.path` and stubs the two plotting / FLOP-count imports that sit on its import path but are not
installed in this image (`matplotlib` — ultralytics/utils/__init__.py:24, `thop` — ultralytics/nn/tasks.py:10).

Search order: $QUAN_REFERENCE_ROOT, <repo>/baseline/_ref (written by baseline/install_ref.py, travels to the GPU box),
/root/reference (the build container's read-only checkout).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types
from pathlib import Path
from typing import Optional

_REPO = Path(__file__).resolve().parent.parent


def find_reference() -> Optional[Path]:
    cands = [os.environ.get("QUAN_REFERENCE_ROOT"), _REPO / "baseline" / "_ref", "/root/reference"]
    for c in cands:
        if c and (Path(c) / "ultralytics" / "nn" / "modules" / "conv.py").exists():
            return Path(c)
    return None


def _stub_missing():
    def have(name):
        try:
            __import__(name)
            return True
        except Exception:                                         # noqa: BLE001
            return False

    if not have("matplotlib"):
        m = types.ModuleType("matplotlib")
        m.use = lambda *a, **k: None
        m.rc = lambda *a, **k: None
        m.rcParams = {}
        m.__path__ = []
        pp = types.ModuleType("matplotlib.pyplot")
        m.pyplot = pp
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = m, pp
    if not have("thop"):
        sys.modules["thop"] = types.ModuleType("thop")


_active = {}


def activate(root: Optional[Path] = None) -> Path:
    """Make `import ultralytics`, `import quaternion.qconv`, `import models.quaternion_models` resolve to the reference."""
    root = Path(root) if root else find_reference()
    if root is None:
        raise RuntimeError("no reference tree found: run `python baseline/install_ref.py` in the build container "
                           "or set QUAN_REFERENCE_ROOT")
    if _active.get("root") == root:
        return root
    os.environ.setdefault("YOLO_CONFIG_DIR", tempfile.mkdtemp(prefix="yolo_cfg_"))   # the package writes settings.json on import
    os.environ.setdefault("YOLO_OFFLINE", "1")
    sys.dont_write_bytecode = True
    _stub_missing()
    for p in (root / "classification", root):
        if str(p) not in sys.path:
            sys.path.insert(0, str(p))
    _active["root"] = root
    return root


def yolo_cfg(name: str = "yolo11n-obb-quan.yaml") -> Path:
    """Path of a model yaml inside the active reference tree (ultralytics/cfg/models/11/)."""
    root = activate()
    p = root / "ultralytics" / "cfg" / "models" / "11" / name
    if not p.exists():      # scale letter in the file name is parsed by the reference (nn/tasks.py:1109-1132); the file has none
        base = name.replace("11n", "11").replace("11s", "11").replace("11m", "11").replace("11l", "11").replace("11x", "11")
        p = root / "ultralytics" / "cfg" / "models" / "11" / base
    return p
