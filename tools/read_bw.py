"""Read-only vs copy HBM bandwidth on this GPU (context for the IQBN reduction rooflines)."""
import torch
def t(fn, it=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / it * 1e3
xs = [torch.randn(64 << 20, device="cuda").to(torch.bfloat16) for _ in range(3)]      # 134 MB each
big = torch.randn(256 << 20, device="cuda").to(torch.bfloat16)                          # 537 MB
out = torch.empty_like(xs[0])
i = [0]
def rot():
    i[0] = (i[0] + 1) % 3
    return xs[i[0]]
S = xs[0].numel() * 2
for name, fn, b in [("sum(bf16, 134MB)", lambda: rot().sum(dtype=torch.float32), S),
                    ("amax(bf16, 134MB)", lambda: rot().amax(), S),
                    ("sum(bf16, 537MB)", lambda: big.sum(dtype=torch.float32), big.numel() * 2),
                    ("copy(134MB)", lambda: out.copy_(rot()), 2 * S)]:
    us = t(fn)
    print(f"{name}: {us:.1f} us  {b / us / 1e3:.0f} GB/s")
