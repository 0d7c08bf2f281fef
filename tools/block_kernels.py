"""Kernel-by-kernel device time of one `Conv` block (QConv2D -> IQBN -> SiLU) forward + backward, for chosen layer shapes of the model
trace (torch.profiler / CUPTI durations; the host is not in the measurement).
    python tools/block_kernels.py "4,8,3,2,1,512" "16,16,3,1,16,128"      # Ci,Co,k,s,groups,H_in  (quaternion channels)"""
import sys
from collections import OrderedDict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import quan_ultralytics_b200 as Q  # noqa: E402


def main():
    from torch.profiler import ProfilerActivity, profile
    B = 16
    for spec in sys.argv[1:]:
        ci, co, k, s, g, H = (int(v) for v in spec.split(","))
        blk = (Q.DWConv(ci * 4, co * 4, k, s) if g > 1 else Q.Conv(ci * 4, co * 4, k, s)).cuda().train()
        x = torch.randn(B, ci, H, H, 4, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last_3d).requires_grad_(True)

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = blk(x)
            y.backward(torch.ones_like(y))

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step()
            torch.cuda.synchronize()
        ks = OrderedDict()
        for e in prof.events():
            for kk in getattr(e, "kernels", None) or []:
                ks.setdefault(kk.name[:110], []).append(kk.duration)
        xb = x.numel() * 2 / 1e6
        print(f"== Conv({ci}->{co}, k{k}, s{s}, g{g}) on {B}x{ci}x{H}^2: input {xb:.0f} MB, output {xb * co / ci / s / s:.0f} MB")
        for n, d in ks.items():
            print(f"   {sum(d):8.1f} us  x{len(d)}  {n}")
        print(f"   total {sum(sum(d) for d in ks.values()):.1f} us")


if __name__ == "__main__":
    main()
