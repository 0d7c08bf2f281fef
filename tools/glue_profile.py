"""Which torch ops (the reference's own Python between our kernels) cost device time in one eager QUAN-YOLO11n-OBB training step?
Kernels that are NOT the library's, attributed to the autograd node / aten op that launched them and to that op's input shapes.
Usage (GPU box): python tools/glue_profile.py [--batch 16] [--size 1024] [--top 60]"""
import argparse
import sys
from collections import defaultdict
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from quan_ultralytics_b200 import loss as qloss, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--top", type=int, default=60)
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    model = workloads.build_yolo_obb("n", 15, dev, swapped=True).train()
    batch = workloads.synthetic_obb_batch(a.batch, a.size, dev)
    tg, tm = qloss.pad_targets(batch, a.batch)
    tg, tm = tg.to(dev), tm.to(dev)
    crit = qloss.OBBLossFused(model)

    def step():
        for p in model.parameters():
            p.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            preds = model(batch["img"])
            l = crit(preds, {"targets": tg, "target_mask": tm})[0]
        l.sum().backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        step()
        torch.cuda.synchronize()
    by = defaultdict(lambda: [0.0, 0])
    lib_us = other_us = 0.0
    for e in prof.events():
        ks = getattr(e, "kernels", None)
        if not ks:
            continue
        top = e
        chain = [e.name]
        while top.cpu_parent is not None:
            top = top.cpu_parent
            chain.append(top.name)
        node = next((n for n in chain if n.startswith("autograd::engine::evaluate_function")), chain[-1])
        node = node.replace("autograd::engine::evaluate_function: ", "bwd ")
        for k in ks:
            if "quan::" in k.name:
                lib_us += k.duration
                continue
            other_us += k.duration
            key = (node, e.name, str(e.input_shapes)[:110])
            by[key][0] += k.duration
            by[key][1] += 1
    print(f"library kernels {lib_us / 1e3:.2f} ms, other kernels {other_us / 1e3:.2f} ms in one eager step")
    agg = defaultdict(lambda: [0.0, 0])
    for (node, op, shp), (us, n) in by.items():
        agg[(node, op)][0] += us
        agg[(node, op)][1] += n
    print("-- by (autograd node | top-level op, launching op)")
    for (node, op), (us, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:30]:
        print(f"{us / 1e3:8.3f} ms {n:5d} x  {node[:44]:44s} {op}")
    print("-- with input shapes")
    for (node, op, shp), (us, n) in sorted(by.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f"{us / 1e3:8.3f} ms {n:4d} x  {node[:34]:34s} {op[:22]:22s} {shp}")


if __name__ == "__main__":
    main()
