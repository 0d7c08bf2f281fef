"""Where does the QUAN-YOLO11n-OBB training step (BASELINE config[2]) spend its time on the B200?  Phases (forward / loss / backward /
clip+SGD) by CUDA events and host wall clock, library launches per step, and a torch.profiler kernel table (library kernels AND the
glue kernels of the reference's own Python: cat / chunk copies, attention matmuls, loss).  Usage (GPU box):
    python tools/yolo_step_profile.py [--batch 16] [--size 1024] [--scale n] [--ref] [--dtype bf16|f32] [--table 45]
"""
import argparse
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from quan_ultralytics_b200 import workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--scale", default="n")
    ap.add_argument("--ref", action="store_true", help="the untouched reference (PyTorch-GPU path) instead of the swapped model")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--table", type=int, default=45)
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    model = workloads.build_yolo_obb(a.scale, 15, dev, swapped=not a.ref).train()
    opt = workloads.yolo_sgd(model)
    batch = workloads.synthetic_obb_batch(a.batch, a.size, dev)
    ac = torch.bfloat16 if a.dtype == "bf16" else None
    import quan_ultralytics_b200 as Q
    lib = Q._lib.load()

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def step(rec=None):
        t = [time.perf_counter()]
        e = [ev()]
        with torch.autocast("cuda", dtype=ac, enabled=ac is not None):
            preds = model(batch["img"])
        e.append(ev()); t.append(time.perf_counter())
        with torch.autocast("cuda", dtype=ac, enabled=ac is not None):
            loss, items = model.loss(batch, preds)
        e.append(ev()); t.append(time.perf_counter())
        loss.backward()
        e.append(ev()); t.append(time.perf_counter())
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=10.0)
        opt.step()
        opt.zero_grad(set_to_none=True)
        e.append(ev()); t.append(time.perf_counter())
        if rec is not None:
            torch.cuda.synchronize()
            rec.append(([e[i].elapsed_time(e[i + 1]) for i in range(4)], [1e3 * (t[i + 1] - t[i]) for i in range(4)], float(loss.detach())))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    n0 = lib.quan_launch_count()
    rec = []
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step(rec)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.steps * 1e3
    names = ("forward", "loss", "backward", "clip+sgd")
    print(f"model: yolo11{a.scale}-obb-quan {'REFERENCE (unswapped)' if a.ref else 'B200 modules'} batch {a.batch} @ {a.size}^2 {a.dtype}")
    print(f"library launches / step: {(lib.quan_launch_count() - n0) / a.steps:.0f};  wall {wall:.1f} ms/step = {a.batch / wall * 1e3:.1f} img/s;  loss {rec[-1][2]:.4f}")
    for i, n in enumerate(names):
        d = sorted(r[0][i] for r in rec)[len(rec) // 2]
        h = sorted(r[1][i] for r in rec)[len(rec) // 2]
        print(f"  {n:9s} device-span {d:8.2f} ms   host {h:8.2f} ms")
    if a.table:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        from torch.autograd import DeviceType
        rows = []
        for e in prof.key_averages():
            if e.device_type != DeviceType.CUDA:
                continue
            t = getattr(e, "self_device_time_total", None)
            if t is None:
                t = e.self_cuda_time_total
            rows.append((t, e.count, e.key))
        rows.sort(reverse=True)
        tot = sum(r[0] for r in rows)
        ours = sum(r[0] for r in rows if "quan::" in r[2])
        print(f"device kernels in one step: {tot / 1e3:.2f} ms in {sum(r[1] for r in rows)} launches; library (quan::) {ours / 1e3:.2f} ms in "
              f"{sum(r[1] for r in rows if 'quan::' in r[2])} launches; everything else {(tot - ours) / 1e3:.2f} ms")
        for t, c, k in rows[: a.table]:
            print(f"{t / 1e3:8.3f} ms {100 * t / tot:5.1f}% {c:5d} x {t / max(c, 1):8.1f} us  {k[:150]}")


if __name__ == "__main__":
    main()
