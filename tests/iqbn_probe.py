"""Tuning probe for the IQBN kernels: times each op on the bench shape for the QUAN_IQBN_* overrides in the env."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import quan_ultralytics_b200 as Q
from quan_ultralytics_b200 import ops

def t(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

N, C, H = int(os.environ.get("N", 64)), int(os.environ.get("C", 256)), int(os.environ.get("HW", 32))
dtype = torch.bfloat16 if os.environ.get("DT", "bf16") == "bf16" else torch.float32
L = ops.LAYOUT_BHWQC
dev = "cuda:0"
xs = [torch.randn(N, C, H, H, 4, device=dev).to(dtype).contiguous(memory_format=torch.channels_last_3d) for _ in range(3)]
dys = [torch.randn_like(x) for x in xs]
outs = [torch.empty_like(x) for x in xs]
gamma, beta = torch.ones(C, 4, device=dev), torch.zeros(C, 4, device=dev)
S = xs[0].numel() * xs[0].element_size()
cnt = float(N * H * H)
stats = ops.iqbn_train_stats(xs[0], L, gamma, beta, 1e-5, 0.1, None, None)
sums = ops.iqbn_bwd_reduce(dys[0], xs[0], L, stats, gamma, beta, Q.ACT_SILU, cnt)
i = [0]
def rot():
    i[0] = (i[0] + 1) % 3
    return i[0]
res = {}
res["stats"] = (t(lambda: ops.iqbn_train_stats(xs[rot()], L, gamma, beta, 1e-5, 0.1, None, None)), 1)
res["apply"] = (t(lambda: ops.iqbn_apply_fwd(xs[rot()], L, stats, gamma, beta, Q.ACT_SILU)), 2)
res["apply_noact"] = (t(lambda: ops.iqbn_apply_fwd(xs[rot()], L, stats, gamma, beta, Q.ACT_NONE)), 2)
res["bwd_reduce"] = (t(lambda: ops.iqbn_bwd_reduce(dys[rot()], xs[i[0]], L, stats, gamma, beta, Q.ACT_SILU, cnt)), 2)
res["bwd_apply"] = (t(lambda: ops.iqbn_bwd_apply(dys[rot()], xs[i[0]], L, stats, gamma, beta, Q.ACT_SILU, sums, cnt)), 3)
res["copy(torch)"] = (t(lambda: outs[rot()].copy_(xs[i[0]])), 2)
tag = f"BPS={os.environ.get('QUAN_IQBN_BPS','-')} U={os.environ.get('QUAN_IQBN_U','-')}"
print(tag, " ".join(f"{k}={v[0]:.1f}us/{v[1]*S/v[0]/1e3:.0f}GB/s" for k, v in res.items()), flush=True)
