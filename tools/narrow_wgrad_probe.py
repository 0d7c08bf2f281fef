"""Device time of the weight-gradient pass alone for the narrow / depthwise layers of the model traces (CUDA-graphed,
10 launches per replay): python tools/narrow_wgrad_probe.py   (GPU box; env toggles select kernel forms)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from quan_ultralytics_b200 import ops  # noqa: E402

# (B, C_i, C_o, k, s, groups, H_o)
SHAPES = [(16, 1, 4, 3, 2, 1, 512), (16, 4, 2, 3, 1, 1, 256), (16, 2, 4, 3, 1, 1, 256), (8, 1, 8, 3, 2, 1, 512),
          (256, 1, 16, 7, 2, 1, 112), (16, 16, 16, 3, 1, 16, 128), (16, 32, 32, 3, 1, 32, 64), (16, 64, 64, 3, 1, 64, 32)]


def main():
    dev = torch.device("cuda", 0)
    for B, ci, co, k, s, g, ho in SHAPES:
        hin = ho * s
        x = torch.randn(B, ci, hin, hin, 4, device=dev).bfloat16().contiguous(memory_format=torch.channels_last_3d)
        dy = torch.randn(B, co, ho, ho, 4, device=dev).bfloat16().contiguous(memory_format=torch.channels_last_3d)
        w = [torch.randn(co, ci // g, k, k, device=dev) for _ in range(4)]
        fn = lambda: ops.qconv2d_bwd(dy, x, w, (s, s), (k // 2, k // 2), (1, 1), g, ops.M_A, False, True, False)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(10):
                fn()
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            graph.replay()
        e1.record()
        e1.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        mb = (x.numel() + dy.numel()) * 2 / 1e6
        print(f"({B},{ci}->{co},k{k},s{s},g{g},{ho}^2): wgrad {us:8.1f} us  ({mb:7.1f} MB in: {mb / us * 1e-3 * 1e3:6.2f} TB/s)", flush=True)


if __name__ == "__main__":
    main()
