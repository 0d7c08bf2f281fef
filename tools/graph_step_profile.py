"""Kernel timeline of the CUDA-graphed QUAN-YOLO11n-OBB training step (bench.py's default workload): per-kernel device time summed by
name and by family (library conv / IQBN / attention / loss+assigner / optimizer / torch glue), GPU-busy share of the step (union of
kernel intervals vs step span) from a torch.profiler (CUPTI) trace of graph replays.
    python tools/graph_step_profile.py [--batch 16] [--size 1024] [--top 40]"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from quan_ultralytics_b200 import optim, workloads  # noqa: E402
from quan_ultralytics_b200.graphs import GraphedTrainStep  # noqa: E402
from quan_ultralytics_b200.loss import OBBLossFused, OBBLossStatic, pad_targets  # noqa: E402


def family(name):
    n = name
    if "quan::" in n:
        for key, fam in (("qconv", "lib conv"), ("pack_weights", "lib conv aux"), ("wgrad_reduce", "lib conv aux"), ("mix", "lib conv aux"),
                         ("iqbn_fold", "lib iqbn fold"), ("iqbn", "lib iqbn"), ("qattn", "lib attention"), ("tal_", "lib assigner"), ("sgd_", "lib optimizer"),
                         ("qmaxpool", "lib pool"), ("upsample", "lib upsample"), ("poincare", "lib poincare"), ("layout", "lib layout")):
            if key in n:
                return fam
        return "lib other"
    if "Memcpy" in n or "Memset" in n:
        return "memcpy/memset"
    if "cudnn" in n or "cutlass" in n or "gemm" in n.lower() or "nchw" in n.lower() or "conv" in n.lower():
        return "torch conv/gemm (QER)"
    return "torch elementwise/reduce (glue + loss)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    torch.manual_seed(0)
    model = workloads.build_yolo_obb("n", 15, "cuda", swapped=True).train()
    opt = optim.yolo_clip_sgd(model)
    batch = workloads.synthetic_obb_batch(a.batch, a.size, "cuda")
    crit = OBBLossFused(model)
    tg, tm = pad_targets(batch, a.batch)
    gs = GraphedTrainStep(lambda img, t, m: model(img), lambda preds, img, t, m: crit(preds, {"targets": t, "target_mask": m}), opt,
                          [batch["img"], tg.cuda(), tm.cuda()], list(model.parameters()), autocast=torch.bfloat16, capture_loss=True)
    for _ in range(3):
        gs(gs.static_inputs)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    reps = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            gs(gs.static_inputs)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    ivs = sorted((e.time_range.start, e.time_range.end, e.name) for e in evs)
    span = ivs[-1][1] - ivs[0][0]
    busy, cur_s, cur_e = 0.0, None, None
    for s, e, _ in ivs:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    by_name, by_fam = {}, {}
    for s, e, n in ivs:
        d = e - s
        t = by_name.setdefault(n, [0.0, 0])
        t[0] += d; t[1] += 1
        f = by_fam.setdefault(family(n), [0.0, 0])
        f[0] += d; f[1] += 1
    print(f"step span {span / reps / 1e3:.2f} ms; GPU busy (union of kernel intervals) {busy / reps / 1e3:.2f} ms = {100 * busy / span:.1f}%; "
          f"{len(ivs) / reps:.0f} device activities per step; summed kernel time {sum(v[0] for v in by_name.values()) / reps / 1e3:.2f} ms")
    for f, (t, c) in sorted(by_fam.items(), key=lambda kv: -kv[1][0]):
        print(f"  {t / reps / 1e3:7.3f} ms  {c / reps:6.0f} launches  {f}")
    print("top kernels:")
    for n, (t, c) in sorted(by_name.items(), key=lambda kv: -kv[1][0])[: a.top]:
        print(f"  {t / reps / 1e3:7.3f} ms {c / reps:6.0f} x {t / c:7.1f} us  {n[:140]}")


if __name__ == "__main__":
    main()
