"""QUAN-YOLO11n-OBB training step (BASELINE config[2]) through GraphedTrainStep: ms / step, device spans of its three phases.
    python tools/yolo_graph_step.py [--batch 16] [--size 1024] [--steps 10]"""
import argparse
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from quan_ultralytics_b200 import optim, workloads  # noqa: E402
from quan_ultralytics_b200.graphs import GraphedTrainStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--scale", default="n")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--nvtx", action="store_true", help="wrap ONE replayed step in the NVTX range 'quanstep' (ncu --nvtx --nvtx-include quanstep/)")
    ap.add_argument("--static-loss", action="store_true", help="loss.OBBLossStatic captured inside the forward graph (whole step = two replays)")
    a = ap.parse_args()
    torch.manual_seed(0)
    model = workloads.build_yolo_obb(a.scale, 15, "cuda", swapped=True).train()
    opt = optim.yolo_clip_sgd(model)
    batch = workloads.synthetic_obb_batch(a.batch, a.size, "cuda")
    t0 = time.perf_counter()
    if a.static_loss:
        from quan_ultralytics_b200.loss import OBBLossFused, OBBLossStatic, pad_targets
        crit = OBBLossFused(model)
        tg, tm = pad_targets(batch, a.batch)
        step = GraphedTrainStep(lambda img, t, m: model(img), lambda preds, img, t, m: crit(preds, {"targets": t, "target_mask": m}), opt,
                                [batch["img"], tg.cuda(), tm.cuda()], list(model.parameters()), autocast=torch.bfloat16, capture_loss=True)
        inputs, largs = [batch["img"], tg.cuda(), tm.cuda()], ()
    else:
        step = GraphedTrainStep(lambda img: model(img), lambda preds, img, b: model.loss(b, preds), opt, [batch["img"]],
                                list(model.parameters()), autocast=torch.bfloat16, loss_args=(batch,))
        inputs, largs = [batch["img"]], (batch,)
    torch.cuda.synchronize()
    print(f"capture: {time.perf_counter() - t0:.1f} s")
    for _ in range(3):
        step(inputs, largs)
    torch.cuda.synchronize()
    if a.nvtx:
        torch.cuda.nvtx.range_push("quanstep")
        step(inputs, largs)
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
        return
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    t0 = time.perf_counter()
    ev[0].record()
    losses = []
    for i in range(a.steps):
        loss, _ = step(inputs, largs)
        losses.append(loss.detach().clone())
        ev[i + 1].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.steps * 1e3
    dev = ev[0].elapsed_time(ev[-1]) / a.steps
    print(f"graphed step: device {dev:.2f} ms, wall {wall:.2f} ms/step = {a.batch / wall * 1e3:.0f} img/s; losses {[round(float(l), 3) for l in losses]}")
    if a.static_loss:
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record(); step.g_fwd.replay(); e[1].record(); step.g_bwd.replay(); e[2].record()
        torch.cuda.synchronize()
        print(f"phases: forward+loss graph {e[0].elapsed_time(e[1]):.2f} ms, backward+optimizer graph {e[1].elapsed_time(e[2]):.2f} ms")
        return
    # phases
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(); step.g_fwd.replay(); e[1].record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    leaves = [t.detach().requires_grad_(t.requires_grad) for t in step.out_flat]
    from torch.utils._pytree import tree_unflatten
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss, _ = model.loss(batch, tree_unflatten(leaves, step.out_spec))
    grads = torch.autograd.grad(loss, [l for l in leaves if l.requires_grad])
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    e[2].record(); step.g_bwd.replay(); e[3].record()
    torch.cuda.synchronize()
    print(f"phases: forward graph {e[0].elapsed_time(e[1]):.2f} ms, loss fwd+bwd (eager, host-synchronous) {1e3 * (t2 - t1):.2f} ms, "
          f"backward+optimizer graph {e[2].elapsed_time(e[3]):.2f} ms")


if __name__ == "__main__":
    main()
