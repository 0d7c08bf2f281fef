"""Golden vectors for QER from the REAL reference (build container only: needs /root/reference):
    python tests/golden/make_qer_golden.py        # writes tests/golden/qer.npz
ultralytics/nn/modules/head.py:26-47 QER(in_channels, out_channels, 1) — forward and autograd backward in fp64 for the three
extractions of the QUAN-YOLO11n OBB head (head.py:114,123,335: 64 -> 64 box, 64 -> 15 class, 16 -> 1 angle) on small ragged maps
(pixel counts that are not multiples of the kernels' 128-pixel tile).
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import import_reference, t2n  # noqa: E402

OUT = Path(__file__).resolve().parent
CASES = {"box_64_64": (64, 64, (2, 9, 11)), "cls_64_15": (64, 15, (3, 7, 10)), "angle_16_1": (16, 1, (2, 13, 9)), "s_cls_128_15": (128, 15, (1, 12, 12))}


def main():
    import_reference()
    import ultralytics.nn.modules.head as uhead
    rec = {}
    torch.manual_seed(11)
    for name, (cin, cout, (B, H, W)) in CASES.items():
        m = uhead.QER(cin, cout, 1).double()
        x = torch.randn(B, cin // 4, H, W, 4, dtype=torch.float64, requires_grad=True)
        y = m(x)
        dy = torch.randn_like(y)
        y.backward(dy)
        rec[f"{name}/x"], rec[f"{name}/w"], rec[f"{name}/b"] = t2n(x), t2n(m.output_proj.weight), t2n(m.output_proj.bias)
        rec[f"{name}/y"], rec[f"{name}/dy"], rec[f"{name}/dx"] = t2n(y), t2n(dy), t2n(x.grad)
        rec[f"{name}/dw"], rec[f"{name}/db"] = t2n(m.output_proj.weight.grad), t2n(m.output_proj.bias.grad)
    np.savez_compressed(OUT / "qer.npz", **rec)
    print("wrote", OUT / "qer.npz", {k: v.shape for k, v in rec.items()})


if __name__ == "__main__":
    main()
