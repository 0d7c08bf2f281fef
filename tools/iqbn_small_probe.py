"""Device time of the IQBN backward reduction (+ its fold, when it has one) per activation shape, graph-replayed (no host in the loop).
Run once per setting: QUAN_IQBN_SMALL_MB=0 / 40 python tools/iqbn_small_probe.py"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import quan_ultralytics_b200 as Q  # noqa: E402
from quan_ultralytics_b200 import ops  # noqa: E402


def gtime(fn, reps=8, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    e1.synchronize()
    return 1e3 * e0.elapsed_time(e1) / (iters * reps)


print("QUAN_IQBN_SMALL_MB =", os.environ.get("QUAN_IQBN_SMALL_MB", "(default)"))
L = ops.LAYOUT_BHWQC
for (C, H) in ((16, 32), (32, 32), (64, 32), (8, 64), (16, 64), (32, 64), (4, 128), (16, 128), (8, 256)):
    x = torch.randn(16, C, H, H, 4, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last_3d)
    dy = torch.randn_like(x)
    gamma, beta = torch.ones(C, 4, device="cuda"), torch.zeros(C, 4, device="cuda")
    stats = ops.iqbn_train_stats(x, L, gamma, beta, 1e-5, 0.1, None, None)
    cnt = float(16 * H * H)
    t_b = gtime(lambda: ops.iqbn_bwd_reduce(dy, x, L, stats, gamma, beta, Q.ACT_SILU, cnt))
    t_f = gtime(lambda: ops.iqbn_train_stats(x, L, gamma, beta, 1e-5, 0.1, None, None))
    mb = x.numel() * 2 / 1e6
    print(f"C={C:3d} {H:3d}^2  tensor {mb:6.1f} MB   bwd reduce {t_b:6.1f} us   fwd stats {t_f:6.1f} us")
