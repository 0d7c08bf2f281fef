"""Data-parallel helpers: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch) for the plumbing.

The hot path shards by batch: weights are replicated, every QConv2D / QUpsample / Poincare output
depends only on its own image.  The two cross-sample couplings are the weight gradients (DDP bucketed all-reduce,
overlapped with backward) and — optionally — IQBN batch statistics ("synced IQBN", new work: the reference has no
cross-rank statistics).  Synced IQBN all-reduces [8C] fp64 sums per layer in forward and again in
backward: latency-bound messages <= 32 KB (the exchange itself: functional._IQBNTrain, `dist.all_reduce` of the fp64 sums, captured into the step graphs like any
other collective).

  convert_sync_iqbn(model)            mark every IQBN to reduce its statistics across the group
  BucketedGradSync(params, nbuckets)  gradient averaging inside a captured (or eager) backward: a bucket's all-reduce is forked
                                      onto a communication stream as soon as its last gradient has arrived (the reference's DDP:
                                      engine/trainer.py:273), joined before the optimizer
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from .modules import IQBN


def convert_sync_iqbn(module: nn.Module, process_group=None) -> nn.Module:
    """Mark every IQBN in `module` to reduce its batch statistics across `process_group` (default: WORLD).
    Mirrors torch.nn.SyncBatchNorm.convert_sync_batchnorm; parameters and buffers are untouched."""
    for m in module.modules():
        if isinstance(m, IQBN):
            m.sync = True
            m.process_group = process_group
    return module


class BucketedGradSync:
    """Data-parallel gradient averaging for a captured (or eager) backward.  The parameters that receive gradients are split, in
    reverse registration order (roughly the order backward produces them), into `nbuckets` buckets of similar size.  Every parameter
    carries a post-accumulate hook; when the LAST gradient of a bucket has arrived — whatever order autograd delivers them in — the
    bucket is packed into its flat buffer, all-reduced (NCCL, average) on `comm_stream` and scattered back, overlapping the rest of
    backward; `finish()` joins the communication stream before the optimizer reads the gradients."""

    def __init__(self, params: Sequence[torch.nn.Parameter], nbuckets: int = 3, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.nbuckets = nbuckets
        self.params = [p for p in params if p.requires_grad]
        self.comm_stream = torch.cuda.Stream(device=self.params[0].device)
        self.buckets: List[List[torch.nn.Parameter]] = []
        self._pending, self._launched, self._handles, self._count = [], set(), [], []
        self.prepare(self.params)

    def prepare(self, used: Sequence[torch.nn.Parameter]) -> None:
        """(Re)build the buckets from the parameters that actually receive gradients (GraphedTrainStep passes the ones its warm-up
        backward touched: a parameter that never gets a gradient would keep its bucket from completing)."""
        keep = set(id(p) for p in used)
        ps = [p for p in reversed(self.params) if id(p) in keep]
        total = sum(p.numel() for p in ps)
        target = max(1, total // max(1, self.nbuckets))
        self.buckets, acc = [[]], 0
        for p in ps:
            self.buckets[-1].append(p)
            acc += p.numel()
            if acc >= target and len(self.buckets) < self.nbuckets:
                self.buckets.append([])
                acc = 0
        self.buckets = [b for b in self.buckets if b]
        dev = ps[0].device
        self.flat = [torch.zeros(sum(p.numel() for p in b), dtype=ps[0].dtype, device=dev) for b in self.buckets]
        self.views = [list(torch._utils._unflatten_dense_tensors(f, b)) for f, b in zip(self.flat, self.buckets)]
        self._count = [0] * len(self.buckets)

    def attach(self) -> None:
        self.detach()
        for bi, bucket in enumerate(self.buckets):
            for p in bucket:
                self._handles.append(p.register_post_accumulate_grad_hook(lambda _p, bi=bi: self._arrived(bi)))

    def detach(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []

    def _arrived(self, bi: int) -> None:
        self._count[bi] += 1
        if self._count[bi] == len(self.buckets[bi]):
            self._launch(bi)

    def _launch(self, bi: int) -> None:
        if bi in self._launched:
            return
        self._launched.add(bi)
        pairs = [(p.grad, v) for p, v in zip(self.buckets[bi], self.views[bi]) if p.grad is not None]
        if not pairs:
            return
        grads, views = [g for g, _ in pairs], [v for _, v in pairs]
        from . import ops
        ops.deferred_wgrad_join(grads[0].device)       # weight gradients forked onto the side stream (ops.deferred_wgrad) land first
        ev = torch.cuda.Event()
        ev.record()                                    # every gradient of the bucket is complete on the backward stream here
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            torch._foreach_copy_(views, grads)         # pack
            dist.all_reduce(self.flat[bi], op=dist.ReduceOp.AVG, group=self.group)
            torch._foreach_copy_(grads, views)         # scatter the averages back where the optimizer reads them
            done = torch.cuda.Event()
            done.record()
        self._pending.append(done)

    def finish(self) -> None:
        """Join the communication stream (call after backward, before the optimizer); a bucket that never completed (hooks not
        attached, or a parameter without gradient this step) is reduced now."""
        for bi in range(len(self.buckets)):
            self._launch(bi)
        for ev in self._pending:
            torch.cuda.current_stream().wait_event(ev)
        self._pending, self._launched = [], set()
        self._count = [0] * len(self.buckets)
