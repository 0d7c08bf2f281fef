"""Timing probe for the QConv2D engines on one sweep shape (env: N, C, HW, DT, S=stride, K=kernel)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from quan_ultralytics_b200 import ops

def t(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

N, C, H = int(os.environ.get("N", 64)), int(os.environ.get("C", 256)), int(os.environ.get("HW", 32))
K, S = int(os.environ.get("K", 3)), int(os.environ.get("S", 1))
dtype = torch.bfloat16 if os.environ.get("DT", "bf16") == "bf16" else torch.float32
L = ops.LAYOUT_BHWQC
dev = "cuda:0"
x = torch.randn(N, C, H, H, 4, device=dev).to(dtype).contiguous(memory_format=torch.channels_last_3d)
w = [torch.randn(C, C, K, K, device=dev) * 0.02 for _ in range(4)]
args = ((S, S), (K // 2, K // 2), (1, 1), 1, ops.M_A)
y = ops.qconv2d_fwd(x, w, None, *args, ops.ALGO_AUTO, L)
dy = torch.randn_like(y)
flops = 4 * 2 * N * y.shape[2] * y.shape[3] * C * C * K * K
which = os.environ.get("WHICH", "fwd,dx,dw")
out = []
if "fwd" in which:
    us = t(lambda: ops.qconv2d_fwd(x, w, None, *args, ops.ALGO_AUTO, L)); out.append(f"fwd={us:.0f}us/{flops/us/1e6:.0f}TF")
if "dx" in which:
    us = t(lambda: ops.qconv2d_bwd(dy, x, w, *args, True, False, False)); out.append(f"mixT+dgrad={us:.0f}us/{flops/us/1e6:.0f}TF")
if "dw" in which:
    us = t(lambda: ops.qconv2d_bwd(dy, x, w, *args, False, True, False)); out.append(f"mixT+wgrad={us:.0f}us/{flops/us/1e6:.0f}TF")
tag = " ".join(f"{k}={os.environ[k]}" for k in ("QUAN_TC_CG", "QUAN_TC_DBG", "QUAN_TC_STAGES") if k in os.environ)
print(f"[N={N} C={C} HW={H} K={K} S={S} {os.environ.get('DT','bf16')}] {tag} " + " ".join(out), flush=True)
