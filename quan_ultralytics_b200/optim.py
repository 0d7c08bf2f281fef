"""The optimizer half of the training step on the B200: `ClipSGD.step()` = the reference's
`BaseTrainer.optimizer_step` (ultralytics/engine/trainer.py:586-594: clip_grad_norm_(10) -> SGD(nesterov).step() -> zero_grad())
for every parameter tensor in two kernel launches of libquan_sm100.so (`quan_sgd_clip_step`), and `ModelEMA`-style averaging
(ultralytics/utils/torch_utils.py:514-525) in one (`quan_ema_update`).  Host side = bookkeeping only: a chunk table of raw pointers.

Works eagerly (the table is rebuilt when gradient tensors move) and inside CUDA-graph capture (the table upload is a captured
pinned-memory copy; gradient addresses are static across replays).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check

CHUNK = 8192
_CHUNK_DTYPE = np.dtype([("p", "<u8"), ("g", "<u8"), ("off", "<i8"), ("n", "<i4"), ("grp", "<i4")])     # struct quan_opt_chunk


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class _ChunkTable:
    """Device array of quan_opt_chunk + its pinned host mirror (kept alive: a captured graph re-reads it on every replay)."""

    def __init__(self, capacity: int, device):
        self.capacity = capacity
        self.host = torch.empty(capacity * _CHUNK_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        self.dev = torch.empty(capacity * _CHUNK_DTYPE.itemsize, dtype=torch.uint8, device=device)
        self.n = 0
        self.key = None

    def upload(self, rows: np.ndarray, key) -> None:
        assert len(rows) <= self.capacity
        self.host.numpy()[: rows.nbytes] = rows.view(np.uint8).reshape(-1)
        self.dev.copy_(self.host, non_blocking=True)
        self.n, self.key = len(rows), key


def _chunks_of(n: int) -> int:
    return (n + CHUNK - 1) // CHUNK


class ClipSGD:
    """clip_grad_norm_(max_norm) + torch.optim.SGD(momentum, nesterov, dampening, per-group lr / weight_decay).step() (+ zero_grad).

    `param_groups`: list of dicts {"params": [...], "lr": float, "weight_decay": float} (the layout of trainer.py:799-806).
    Parameters and gradients must be dense fp32 CUDA tensors; a parameter whose .grad is None is skipped, as torch does."""

    def __init__(self, param_groups: Sequence[Dict], momentum: float = 0.9, nesterov: bool = True, dampening: float = 0.0,
                 max_norm: float = 10.0):
        self.param_groups = [dict(g) for g in param_groups]
        seen = set()
        for g in self.param_groups:
            g["params"] = [p for p in g["params"] if not (id(p) in seen or seen.add(id(p)))]
            g.setdefault("weight_decay", 0.0)
        params = [p for g in self.param_groups for p in g["params"]]
        if not params:
            raise ValueError("ClipSGD: no parameters")
        self.device = params[0].device
        if self.device.type != "cuda":
            raise RuntimeError("ClipSGD needs CUDA parameters: there is no CPU fallback")
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != self.device:
                raise RuntimeError("ClipSGD: parameters must be dense fp32 tensors on one device")
        self.momentum, self.nesterov, self.dampening, self.max_norm = momentum, nesterov, dampening, max_norm
        self._offsets, total = {}, 0
        for p in params:
            self._offsets[id(p)] = total
            total += (p.numel() + 3) // 4 * 4                      # 16-byte aligned slots
        self.momentum_buf = torch.zeros(total, dtype=torch.float32, device=self.device)
        nchunks = sum(_chunks_of(p.numel()) for p in params)
        self._table = _ChunkTable(nchunks, self.device)
        self._partial = torch.empty(nchunks, dtype=torch.float64, device=self.device)
        self.total_norm = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._hyper_host = torch.empty(2 * len(self.param_groups) + 4, dtype=torch.float32).pin_memory()
        self._hyper = torch.empty_like(self._hyper_host, device=self.device)
        self._hyper_key = None
        self._sync_hyper()

    # -- hyper-parameters live on the device (a captured graph follows schedules): re-upload only when they change
    def _sync_hyper(self) -> None:
        vals = []
        for g in self.param_groups:
            vals += [float(g["lr"]), float(g["weight_decay"])]
        vals += [float(self.momentum), float(self.max_norm or 0.0), 1.0 if self.nesterov else 0.0, float(self.dampening)]
        key = tuple(vals)
        if key != self._hyper_key:
            self._hyper_host.copy_(torch.tensor(vals, dtype=torch.float32))
            self._hyper.copy_(self._hyper_host, non_blocking=True)
            self._hyper_key = key

    def _rows(self):
        rows, key = [], []
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                gr = p.grad
                if gr is None:
                    continue
                if gr.dtype != torch.float32 or not gr.is_contiguous():
                    raise RuntimeError("ClipSGD: gradients must be dense fp32 tensors")
                pp, gp, off, n = p.data_ptr(), gr.data_ptr(), self._offsets[id(p)], p.numel()
                key.append(gp)
                for c in range(0, n, CHUNK):
                    rows.append((pp + 4 * c, gp + 4 * c, off + c, min(CHUNK, n - c), gi))
        return rows, tuple(key)

    def step(self, zero_grad: bool = False) -> None:
        """One optimizer step over every parameter that has a gradient.  zero_grad=True also clears the gradients in the same pass
        (optimizer.zero_grad(set_to_none=False))."""
        self._sync_hyper()
        rows, key = self._rows()
        if key != self._table.key:
            self._table.upload(np.array(rows, dtype=_CHUNK_DTYPE), key)
        t = self._table
        check(_lib.load().quan_sgd_clip_step(t.dev.data_ptr(), t.n, self.momentum_buf.data_ptr(), self._hyper.data_ptr(),
                                             len(self.param_groups), self._partial.data_ptr(), self.total_norm.data_ptr(),
                                             int(zero_grad), _stream(self.device)), "quan_sgd_clip_step")

    def zero_grad(self, set_to_none: bool = True) -> None:
        for g in self.param_groups:
            for p in g["params"]:
                if p.grad is not None:
                    if set_to_none:
                        p.grad = None
                    else:
                        p.grad.zero_()

    def momentum_of(self, p: torch.Tensor) -> torch.Tensor:
        off = self._offsets[id(p)]
        return self.momentum_buf[off:off + p.numel()].view_as(p)


class ParamEMA:
    """ModelEMA.update (torch_utils.py:514-525) over the float entries of a module's state dict: e = d e + (1 - d) v with
    d = decay (1 - exp(-updates / tau)), one launch."""

    def __init__(self, module: torch.nn.Module, decay: float = 0.9999, tau: float = 2000.0, updates: int = 0):
        self.items = [(k, v) for k, v in module.state_dict().items() if v.dtype == torch.float32 and v.is_cuda and v.is_contiguous()]
        if not self.items:
            raise ValueError("ParamEMA: no fp32 CUDA tensors")
        self.device = self.items[0][1].device
        self.decay_fn = lambda x: decay * (1 - math.exp(-x / tau))
        self.updates = updates
        rows, total, self._off = [], 0, {}
        for k, v in self.items:
            self._off[k] = total
            n = v.numel()
            for c in range(0, n, CHUNK):
                rows.append((v.data_ptr() + 4 * c, 0, total + c, min(CHUNK, n - c), 0))
            total += (n + 3) // 4 * 4
        self.ema_buf = torch.zeros(total, dtype=torch.float32, device=self.device)
        for k, v in self.items:
            self.ema_buf[self._off[k]:self._off[k] + v.numel()].copy_(v.detach().reshape(-1))
        self._table = _ChunkTable(len(rows), self.device)
        self._table.upload(np.array(rows, dtype=_CHUNK_DTYPE), None)
        self._d_host = torch.empty(1, dtype=torch.float32).pin_memory()
        self._d = torch.empty(1, dtype=torch.float32, device=self.device)

    def update(self) -> None:
        self.updates += 1
        self._d_host[0] = self.decay_fn(self.updates)
        self._d.copy_(self._d_host, non_blocking=True)
        check(_lib.load().quan_ema_update(self._table.dev.data_ptr(), self._table.n, self.ema_buf.data_ptr(), self._d.data_ptr(),
                                          _stream(self.device)), "quan_ema_update")

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return {k: self.ema_buf[self._off[k]:self._off[k] + v.numel()].view_as(v) for k, v in self.items}


def yolo_clip_sgd(model, lr: float = 0.01, momentum: float = 0.937, decay: float = 5e-4, max_norm: float = 10.0) -> ClipSGD:
    """The reference trainer's optimizer for the YOLO step (trainer.py:766-806 groups, lr0 / momentum / weight_decay of
    cfg/default.yaml, clip 10 of trainer.py:589) as one ClipSGD."""
    from .workloads import yolo_param_groups
    g = yolo_param_groups(model)
    return ClipSGD([{"params": g[2], "lr": lr, "weight_decay": 0.0}, {"params": g[0], "lr": lr, "weight_decay": decay},
                    {"params": g[1], "lr": lr, "weight_decay": 0.0}], momentum=momentum, nesterov=True, max_norm=max_norm)
