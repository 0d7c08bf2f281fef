"""TMA-store epilogue probe: forward / dgrad time of a few QConv2D shapes, each op alone with the L2 swept between launches.
Run once per setting (the switch is read once per process):

    QUAN_TC_TSTORE=0 python tools/tstore_probe.py ; QUAN_TC_TSTORE=1 python tools/tstore_probe.py
"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from bench import time_op  # noqa: E402
from quan_ultralytics_b200 import ops  # noqa: E402

# (label, N, Ci, Co, H, k, s, dtype)
SHAPES = [
    ("sep 1x1 Cq=256 64x32^2", 64, 256, 256, 32, 1, 1, "bf16"),
    ("sep 3x3 Cq=256 64x32^2", 64, 256, 256, 32, 3, 1, "bf16"),
    ("sep 3x3 s2 Cq=256 64x32^2", 64, 256, 256, 32, 3, 2, "bf16"),
    ("sep 3x3 s2 Cq=128 64x64^2", 64, 128, 128, 64, 3, 2, "bf16"),
    ("sep 3x3 Cq=64 64x64^2", 64, 64, 64, 64, 3, 1, "bf16"),
    ("sep 3x3 Cq=256 tf32 64x32^2", 64, 256, 256, 32, 3, 1, "f32"),
    ("dense 1x1 8->8 16x256^2", 16, 8, 8, 256, 1, 1, "bf16"),
    ("dense 1x1 12->16 16x256^2", 16, 12, 16, 256, 1, 1, "bf16"),
    ("dense 3x3 16->16 16x128^2", 16, 16, 16, 128, 3, 1, "bf16"),
    ("dense 3x3 s2 4->8 16x512^2", 16, 4, 8, 512, 3, 2, "bf16"),
    ("dense 1x1 48->32 16x64^2", 16, 48, 32, 64, 1, 1, "bf16"),
]


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L = ops.LAYOUT_BHWQC
    flush = torch.zeros(64 << 20, dtype=torch.int32, device=dev)
    print(f"QUAN_TC_TSTORE={os.environ.get('QUAN_TC_TSTORE', '(default)')}")
    for label, N, Ci, Co, H, k, s, dt in SHAPES:
        dtype = torch.bfloat16 if dt == "bf16" else torch.float32
        torch.manual_seed(0)
        x = torch.randn(N, Ci, H, H, 4, device=dev).to(dtype).contiguous(memory_format=torch.channels_last_3d)
        w = [torch.randn(Co, Ci, k, k, device=dev) / (Ci * k * k) ** 0.5 for _ in range(4)]
        args = ((s, s), (k // 2, k // 2), (1, 1), 1, ops.M_A)
        y = ops.qconv2d_fwd(x, w, None, *args, ops.ALGO_AUTO, L)
        dy = torch.randn_like(y)
        f = time_op(lambda: ops.qconv2d_fwd(x, w, None, *args, ops.ALGO_AUTO, L), 10, flush)
        d = time_op(lambda: ops.qconv2d_bwd(dy, x, w, *args, True, False, False), 10, flush)
        esz = 2 if dt == "bf16" else 4
        mb = (x.numel() + y.numel()) * esz / 1e6
        print(f"{label:32s} fwd {f * 1e3:8.1f} us  dgrad {d * 1e3:8.1f} us   in+out {mb:7.1f} MB  (HBM floor {mb / 6.5517:6.1f} us)")
        del x, y, dy, w
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
