"""Device time of quan_qer_fwd / quan_qer_bwd per head shape (QUAN-YOLO11n OBB head at 16 x 1024^2) against the HBM floor.
Usage (GPU box): python tools/qer_probe.py [--iters 20]"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from quan_ultralytics_b200 import ops  # noqa: E402


def timed(fn, iters, flush):
    """Median device time of one call, taken from a CUDA graph of (L2 flush, call) x 4 minus the flush alone: host launch overhead
    (the Python wrapper costs more than these kernels) stays outside."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()

    def graph_of(with_fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(4):
                flush.zero_()
                if with_fn:
                    fn()
        return g

    def run(g):
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / 4)
        ts.sort()
        return ts[len(ts) // 2]

    return run(graph_of(True)) - run(graph_of(False))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B = 16
    print(f"{'shape':28s} {'pass':8s} {'us':>8s} {'floor us':>9s} {'frac':>6s}")
    for (H, C, N, ld, tag) in [(128, 16, 64, 80, "P3 box"), (128, 16, 15, 80, "P3 cls"), (128, 4, 1, 1, "P3 angle"), (64, 16, 64, 80, "P4 box"),
                               (64, 16, 15, 80, "P4 cls"), (32, 16, 64, 80, "P5 box"), (128, 16, 64, 64, "P3 box dense"), (128, 16, 64, 79, "P3 box ld79")]:
        if a.only and a.only not in tag:
            continue
        x = torch.randn(B, C, H, H, 4, device=dev).bfloat16().contiguous(memory_format=torch.channels_last_3d)
        w = torch.randn(N, 4 * C, 1, 1, device=dev)
        b = torch.randn(N, device=dev)
        buf = torch.empty(B, H, H, ld, device=dev, dtype=torch.bfloat16)
        dy = torch.randn(B, H, H, ld, device=dev).bfloat16()[..., :N].permute(0, 3, 1, 2)
        npix = B * H * H
        by = npix * (4 * C + N) * 2
        floor = by / 6551.7e3
        for name, fn in (("fwd", lambda: ops.qer_fwd(x, w, b, buf, 0, ld if N % 8 else N)), ("dgrad", lambda: ops.qer_bwd(dy, x, w, True, False, False, ld if (N % 8 and ld > 1) else 0)),
                         ("wgrad", lambda: ops.qer_bwd(dy, x, w, False, True, True, ld if (N % 8 and ld > 1) else 0))):
            us = timed(fn, a.iters, flush)
            print(f"{tag + f' K={4*C} N={N} ld={ld}':28s} {name:8s} {us:8.1f} {floor:9.1f} {floor / us:6.2f}")


if __name__ == "__main__":
    main()
