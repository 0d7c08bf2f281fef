"""Multi-GPU (N = 2) on real hardware, collected by `pytest -m gpu` and skipped on a one-GPU box: tests/dist_gpu_check.py under
torchrun — DDP + synced IQBN against the single-process global batch, and the CUDA-graphed data-parallel step (gradient buckets and
IQBN statistics all-reduced inside the captured graphs).  The same protocol runs on CPU / gloo in tests/test_distributed_cpu.py."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_gpu_ddp_synced_iqbn_and_graphed_step():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(ROOT / "tests" / "dist_gpu_check.py")], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
