"""CUDA-graphed training step: the whole forward, and the whole backward + gradient all-reduce + optimizer step, as two graph
replays around the loss (which stays eager while it is the reference's own data-dependent Python — utils/loss.py:941-1033 reads
`.item()`s and indexes with boolean masks — and joins the second graph when the loss function is capturable).

The narrow QUAN models are launch-bound: QUAN-YOLO11n issues ~950 library kernels of 5-30 us per step plus the glue kernels of the
reference's Python (cat / chunk copies, attention, heads); eager mode spends more host time than device time
(tools/yolo_step_profile.py: 43 ms wall for 30 ms of kernels).  Capturing once and replaying removes the host from the step.

    step = GraphedTrainStep(model, loss_fn, optimizer, example_inputs, autocast=torch.bfloat16)
    loss, items = step(img, batch)            # img is copied into the static input buffer; everything else is replay

Mechanics (PyTorch's own whole-network capture recipe, torch.cuda.graphs): warm-up iterations on a side stream, forward captured
with gradients enabled (its autograd graph and saved activations live in the capture's private pool), backward captured into the
same pool with `.grad = None` so that autograd's accumulators adopt pool-resident gradient tensors whose addresses stay fixed;
the optimizer (optim.ClipSGD: chunk table of raw pointers, hyper-parameters in device memory) is captured behind it.  Multi-GPU:
gradients are packed per bucket and all-reduced with NCCL on a side stream INSIDE the captured backward — hooks on the last
gradient of each bucket fork the communication stream, so every replay overlaps bucket k's all-reduce with the rest of backward.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
from torch.utils._pytree import tree_flatten, tree_unflatten

from .distributed import BucketedGradSync  # noqa: F401  (lives with the other data-parallel pieces; re-exported for callers)


class GraphedTrainStep:
    """forward graph -> loss -> backward (+ all-reduce) + optimizer graph.  `forward_fn(*static_inputs)` returns any pytree of
    tensors; `loss_fn(outputs, *static_inputs, *loss_args)` returns (loss, aux) with `loss` a scalar tensor that depends on the
    outputs.  pack_plan=True hoists the weight packing of every tensor-core convolution into one launch at the top of the forward
    graph (ops.PackPlan; the weights only change in the optimizer at the end of the backward graph).  capture_loss=True puts the loss
    into the forward graph (it must then be free of host synchronisation, e.g.
    loss.OBBLossStatic, and read its targets from the static inputs)."""

    def __init__(self, forward_fn: Callable, loss_fn: Callable, optimizer, example_inputs: Sequence[torch.Tensor],
                 params: Sequence[torch.nn.Parameter], autocast: Optional[torch.dtype] = None, warmup: int = 3,
                 loss_args: Sequence = (), grad_sync: Optional[BucketedGradSync] = None, capture_loss: bool = False,
                 pack_plan: bool = True):
        self.forward_fn, self.loss_fn, self.opt = forward_fn, loss_fn, optimizer
        self.params = [p for p in params if p.requires_grad]
        self.autocast = autocast
        self.grad_sync = grad_sync
        self.capture_loss = capture_loss
        self.static_inputs = [torch.empty_like(t).copy_(t) for t in example_inputs]
        dev = self.static_inputs[0].device

        def ac():
            return torch.autocast("cuda", dtype=autocast, enabled=autocast is not None)

        self._ac = ac
        # warm-up on a side stream (workspaces, library handles, allocator state)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        from . import ops
        self.pack_plan = ops.PackPlan() if pack_plan else None
        with torch.cuda.stream(side):
            for i in range(max(1, warmup)):
                if self.pack_plan is not None and i == max(1, warmup) - 1:
                    self.pack_plan.record()                    # note every weight pack the last warm-up step performs
                self._eager(loss_args)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if self.pack_plan is not None and self.pack_plan.commit(dev) == 0:
            self.pack_plan = None                              # nothing on the tensor-core engine: nothing to hoist
        if grad_sync is not None:
            grad_sync.prepare([p for p in self.params if p.grad is not None])     # the parameters this model's backward reaches
        for p in self.params:
            p.grad = None
        from . import _lib
        n0 = _lib.load().quan_launch_count()
        # ---- capture 1: forward (and the loss, when it is capturable)
        self.g_fwd = torch.cuda.CUDAGraph()
        self.loss = self.aux = None
        with torch.cuda.graph(self.g_fwd):
            if self.pack_plan is not None:
                self.pack_plan.run()                           # first node: every layer's packed weights from the current masters
            with ac():
                out = self.forward_fn(*self.static_inputs)
                if capture_loss:
                    self.loss, self.aux = self.loss_fn(out, *self.static_inputs, *loss_args)
        self.out_flat, self.out_spec = tree_flatten(out)
        self.d_out = [torch.zeros_like(t) for t in self.out_flat]
        # ---- capture 2: backward (+ gradient all-reduce) + optimizer, same memory pool
        self.g_bwd = torch.cuda.CUDAGraph()
        if grad_sync is not None:
            grad_sync.attach()
        with torch.cuda.graph(self.g_bwd, pool=self.g_fwd.pool()):
            with ops.deferred_wgrad(dev):                      # wgrad chains off the dX critical path; joined before the optimizer
                if capture_loss:
                    self.loss.backward()
                else:
                    live = [(t, d) for t, d in zip(self.out_flat, self.d_out) if t.requires_grad]
                    torch.autograd.backward([t for t, _ in live], [d for _, d in live])
                if grad_sync is not None:
                    ops.deferred_wgrad_join(dev)
                    grad_sync.finish()
            self.opt.step()
        if grad_sync is not None:
            grad_sync.detach()
        if self.pack_plan is not None:
            self.pack_plan.release()                           # eager calls pack for themselves again; the graphs keep the arena
        self.captured_launches = int(_lib.load().quan_launch_count() - n0)      # library kernels one step replays

    def _eager(self, loss_args):
        for p in self.params:
            p.grad = None
        with self._ac():
            out = self.forward_fn(*self.static_inputs)
            loss, _ = self.loss_fn(out, *self.static_inputs, *loss_args)
        loss.backward()
        # the optimizer is NOT stepped during warm-up: training state starts at the first replay

    def __call__(self, inputs: Sequence[torch.Tensor], loss_args: Sequence = ()):
        for s, t in zip(self.static_inputs, inputs):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.g_fwd.replay()
        if self.capture_loss:
            self.g_bwd.replay()
            return self.loss, self.aux
        leaves = [t.detach().requires_grad_(t.requires_grad) for t in self.out_flat]
        with self._ac():
            loss, aux = self.loss_fn(tree_unflatten(leaves, self.out_spec), *self.static_inputs, *loss_args)
        live = [(l, d) for l, d in zip(leaves, self.d_out) if l.requires_grad]
        grads = torch.autograd.grad(loss, [l for l, _ in live], allow_unused=True)
        for (_, d), g in zip(live, grads):
            if g is None:
                d.zero_()
            else:
                d.copy_(g)
        self.g_bwd.replay()
        return loss, aux
