"""One IQBN backward reduction at a narrow-layer shape, for `ncu --set full -k regex:iqbn_reduce` captures:
    python tools/iqbn_reduce_one.py [C] [H]        (default 16 32: 16 x 16 x 32^2 x 4 bf16 = 2.1 MB)"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import quan_ultralytics_b200 as Q  # noqa: E402
from quan_ultralytics_b200 import ops  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 32
L = ops.LAYOUT_BHWQC
x = torch.randn(16, C, H, H, 4, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last_3d)
dy = torch.randn_like(x)
gamma, beta = torch.ones(C, 4, device="cuda"), torch.zeros(C, 4, device="cuda")
stats = ops.iqbn_train_stats(x, L, gamma, beta, 1e-5, 0.1, None, None)
for _ in range(6):
    ops.iqbn_bwd_reduce(dy, x, L, stats, gamma, beta, Q.ACT_SILU, float(16 * H * H))
torch.cuda.synchronize()
