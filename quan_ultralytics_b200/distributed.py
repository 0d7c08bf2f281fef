"""Data-parallel helpers: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch) for the plumbing.

The hot path shards by batch: weights are replicated, every QConv2D / QUpsample / Poincare output
depends only on its own image.  The two cross-sample couplings are the weight gradients (DDP bucketed all-reduce,
overlapped with backward) and — optionally — IQBN batch statistics ("synced IQBN", new work: the reference has no
cross-rank statistics).  Synced IQBN all-reduces [8C] fp64 sums per layer in forward and again in
backward: latency-bound messages <= 32 KB.
"""
from __future__ import annotations

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
