"""Golden vectors for QuaternionMaxPool from the REAL reference (build container only: needs /root/reference):
    python tests/golden/make_pool_golden.py        # writes tests/golden/qpool.npz
ultralytics/nn/modules/block.py:85-109 QuaternionMaxPool(kernel_size, stride, padding) — forward and autograd backward
in fp64 on inputs quantised to halves (so windows hold ties and the first-maximum rule is exercised), for the two
configurations the models use: QSPPF (5, 1, 2) and the Q-ResNet stem (3, 2, 1), plus the default (2, 2, 0).
The classification copy (classification/models/blocks/quaternion_blocks.py:236-260) is checked to agree.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import import_reference, t2n  # noqa: E402

OUT = Path(__file__).resolve().parent

CASES = {"sppf_k5": (5, 1, 2, (2, 3, 9, 7)), "stem_k3s2": (3, 2, 1, (2, 2, 11, 10)), "default_k2": (2, 2, 0, (1, 4, 8, 6))}


def main():
    import_reference()
    import ultralytics.nn.modules.block as ublock
    from models.blocks.quaternion_blocks import QuaternionMaxPool as CPool
    rec = {}
    g = torch.Generator().manual_seed(7)
    for name, (k, s, p, (B, C, H, W)) in CASES.items():
        x = (torch.randn(B, C, H, W, 4, generator=g, dtype=torch.float64) * 2).round() / 2
        x.requires_grad_(True)
        y = ublock.QuaternionMaxPool(k, s, p)(x)
        dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
        y.backward(dy)
        assert torch.equal(CPool(k, s, p)(x.detach()), y.detach())
        rec[f"{name}/cfg"] = np.array([k, s, p])
        rec[f"{name}/x"], rec[f"{name}/y"], rec[f"{name}/dy"], rec[f"{name}/dx"] = t2n(x), t2n(y), t2n(dy), t2n(x.grad)
    np.savez_compressed(OUT / "qpool.npz", **rec)
    print("wrote", OUT / "qpool.npz", {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    main()
