"""Multi-GPU check, run under torchrun (one process per GPU, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py
Every rank builds the same global batch from a seed, works on its own shard through DDP + synced IQBN, and compares with
the single-process result on the global batch computed locally with the same CUDA kernels."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist

import quan_ultralytics_b200 as Q
from quan_ultralytics_b200.distributed import convert_sync_iqbn


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for dtype, tol in ((torch.float32, 2e-3), (torch.bfloat16, 3e-2)):
        torch.manual_seed(1234)
        Cq, per = 64, 4
        xg = (torch.randn(per * world, Cq, 16, 16, 4, device=dev) * 1.3 + 0.2).to(dtype)
        xg[per:] += 0.5                                         # ranks see different local statistics
        dyg = torch.randn(per * world, Cq, 16, 16, 4, device=dev).to(dtype)

        def make():
            torch.manual_seed(7)
            return torch.nn.Sequential(Q.Conv(Cq * 4, Cq * 4, 3, 1), Q.Conv(Cq * 4, Cq * 4, 3, 1)).to(dev).train()

        # reference: single process, global batch
        ref = make()
        xr = xg.clone().requires_grad_(True)
        yr = ref(xr)
        yr.backward(dyg)
        # distributed: shard + DDP + synced IQBN
        net = convert_sync_iqbn(make())
        ddp = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local])
        xs = xg[rank * per:(rank + 1) * per].clone().requires_grad_(True)
        ys = ddp(xs)
        ys.backward(dyg[rank * per:(rank + 1) * per])
        torch.cuda.synchronize()
        errs = {"y": rel(ys, yr[rank * per:(rank + 1) * per]), "dx": rel(xs.grad, xr.grad[rank * per:(rank + 1) * per])}
        for (n, p), (_, pr) in zip(net.named_parameters(), ref.named_parameters()):
            errs[n] = rel(p.grad * world, pr.grad)             # DDP averages; the global-batch gradient is the sum
        for (n, b), (_, br) in zip(net.named_buffers(), ref.named_buffers()):
            if "running" in n:
                errs[n] = rel(b, br)
        worst = max(errs.values())
        t = torch.tensor([worst], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"[dist check {dtype} world={world}] worst rel err {t.item():.3e} (tol {tol}) "
                  f"y={errs['y']:.2e} dx={errs['dx']:.2e} dW0={errs['0.conv.weight_r']:.2e} dgamma0={errs['0.bn.gamma']:.2e}",
                  flush=True)
        ok &= t.item() <= tol
    ok &= graphed_step_check(rank, world, dev)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)          # (no process-group teardown underneath captured graphs that hold NCCL work)


def graphed_step_check(rank, world, dev):
    """graphs.GraphedTrainStep with the gradient buckets all-reduced INSIDE the captured backward (BucketedGradSync) and the
    synced-IQBN statistics all-reduced inside the captured forward / backward: after one replayed step every rank must hold the
    parameters a single process gets from the global batch (same ClipSGD, gradient = mean over ranks of the shard gradients)."""
    from quan_ultralytics_b200.graphs import BucketedGradSync, GraphedTrainStep
    from quan_ultralytics_b200.optim import ClipSGD
    torch.manual_seed(99)
    Cq, per = 32, 4
    xg = torch.randn(per * world, Cq, 16, 16, 4, device=dev) * 1.3 + 0.2
    xg[per:] += 0.5
    dyg = torch.randn(per * world, Cq, 16, 16, 4, device=dev)

    def make():
        torch.manual_seed(7)
        return torch.nn.Sequential(Q.Conv(Cq * 4, Cq * 4, 3, 1), Q.Conv(Cq * 4, Cq * 4, 1, 1), Q.Conv(Cq * 4, Cq * 4, 3, 1)).to(dev).train()

    def opt_for(net):
        return ClipSGD([{"params": list(net.parameters()), "lr": 0.1, "weight_decay": 1e-3}], momentum=0.9, nesterov=True, max_norm=5.0)

    # single process, global batch; loss = sum(y * dy) / world  ==  mean over ranks of the shard losses
    ref = make()
    oref = opt_for(ref)
    (ref(xg) * dyg).sum().div(world).backward()
    oref.step()
    # data parallel, graphed
    net = convert_sync_iqbn(make())
    params = list(net.parameters())
    sync = BucketedGradSync(params, nbuckets=2)
    sl = slice(rank * per, (rank + 1) * per)
    step = GraphedTrainStep(lambda x, dy: net(x), lambda y, x, dy: ((y * dy).sum(), None), opt_for(net), [xg[sl].clone(), dyg[sl].clone()],
                            params, grad_sync=sync, capture_loss=True)
    with torch.no_grad():                                   # warm-up moved the running statistics; the optimizer has not stepped yet
        for (_, b), (_, b0) in zip(net.named_buffers(), make().named_buffers()):
            b.copy_(b0)
    step(step.static_inputs)
    torch.cuda.synchronize()
    worst = 0.0
    for (n, p), (_, pr) in zip(net.named_parameters(), ref.named_parameters()):
        worst = max(worst, rel(p.detach(), pr.detach()))
    for (n, b), (_, br) in zip(net.named_buffers(), ref.named_buffers()):
        if "running" in n:
            worst = max(worst, rel(b, br))
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"[graphed DP step world={world}] parameters + running statistics after one replay vs single-process global batch: "
              f"worst rel err {t.item():.3e} (tol 2e-3); buckets {[len(b) for b in sync.buckets]}", flush=True)
    return t.item() <= 2e-3


if __name__ == "__main__":
    main()
