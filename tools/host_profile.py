"""Host-side (Python / ctypes / launch) cost of one narrow Conv block step: cProfile over 200 eager steps."""
import cProfile, pstats, sys, io, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import quan_ultralytics_b200 as Q

dev = "cuda:0"
blk = Q.Conv(64, 64, 3, 1).to(dev).train()
x = torch.randn(16, 16, 64, 64, 4, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)
dy = torch.randn(16, 16, 64, 64, 4, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)

def step():
    x.grad = None
    for p in blk.parameters():
        p.grad = None
    y = blk(x)
    y.backward(dy)

for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host time per step {1e6 * (t1 - t0) / 200:.0f} us (device drained after +{1e3 * (t2 - t1):.1f} ms)")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
print("\n".join(l[:150] for l in s.getvalue().splitlines()[:45]))
