"""QuaternionMaxPool device time at the model shapes (GPU box): python tools/pool_probe.py
Algorithmic bytes: fwd = S_in + S_out + idx (1 B per output element), bwd = S_dy + idx + S_dx."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from quan_ultralytics_b200 import ops  # noqa: E402


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


def main():
    # Q-ResNet-34 stem pool (B=256, 16 ch, 112^2 -> 56^2), QSPPF pool of QUAN-YOLO11n / 11s (B=16, 32 / 64 ch, 32^2)
    for B, C, H, k, s, p in [(256, 16, 112, 3, 2, 1), (16, 32, 32, 5, 1, 2), (16, 64, 32, 5, 1, 2)]:
        x = torch.randn(B, C, H, H, 4, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last_3d)
        y, idx = ops.qmaxpool_fwd(x, k, s, p)
        dy = torch.randn_like(y)
        tf = timed(lambda: ops.qmaxpool_fwd(x, k, s, p))
        ti = timed(lambda: ops.qmaxpool_fwd(x, k, s, p, with_idx=False))
        tb = timed(lambda: ops.qmaxpool_bwd(dy, idx, (H, H), k, s, p))
        sx, sy = x.numel() * 2, y.numel() * 2
        print(f"B{B} C{C} {H}^2 k{k}s{s}p{p}: fwd {tf:7.1f} us ({(sx + sy + y.numel()) / tf / 1e3:6.0f} GB/s)  "
              f"fwd(no idx) {ti:7.1f} us ({(sx + sy) / ti / 1e3:6.0f} GB/s)  bwd {tb:7.1f} us "
              f"({(sy + y.numel() + sx) / tb / 1e3:6.0f} GB/s)", flush=True)


if __name__ == "__main__":
    main()
