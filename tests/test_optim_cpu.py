"""The optimizer oracle (oracle/quan_oracle.py sgd_clip_step / ema_update) pinned against the torch calls the reference makes
(engine/trainer.py:586-594 clip_grad_norm_ + SGD.step; utils/torch_utils.py:514-525), and ClipSGD's host-side contract."""
import numpy as np
import pytest
import torch

from oracle import quan_oracle as O


@pytest.mark.parametrize("nesterov,max_norm", [(True, 2.0), (False, 2.0), (True, 0.0)])
def test_oracle_sgd_clip_step_equals_torch(nesterov, max_norm):
    g = torch.Generator().manual_seed(0)
    ps = [torch.randn(5, 3, generator=g, dtype=torch.float64).requires_grad_(True),
          torch.randn(7, generator=g, dtype=torch.float64).requires_grad_(True)]
    P = [p.detach().numpy().copy() for p in ps]
    B = [np.zeros_like(p) for p in P]
    opt = torch.optim.SGD([{"params": [ps[0]], "lr": 0.01, "weight_decay": 1e-3}, {"params": [ps[1]], "lr": 0.02, "weight_decay": 0}],
                          momentum=0.9, nesterov=nesterov)
    for _ in range(4):
        gr = [torch.randn(p.shape, generator=g, dtype=torch.float64) * 3 for p in ps]
        for p, x in zip(ps, gr):
            p.grad = x.clone()
        if max_norm > 0:
            tn = float(torch.nn.utils.clip_grad_norm_(ps, max_norm))
        opt.step()
        P, B, G, total = O.sgd_clip_step(P, [x.numpy() for x in gr], B, [0, 1], [0.01, 0.02], [1e-3, 0.0], 0.9, max_norm, nesterov)
        if max_norm > 0:
            assert abs(tn - total) < 1e-12
        for p, q, gc in zip(ps, P, G):
            np.testing.assert_allclose(p.detach().numpy(), q, rtol=0, atol=1e-14)
            np.testing.assert_allclose(p.grad.numpy(), gc, rtol=0, atol=1e-14)


def test_clip_sgd_refuses_cpu_parameters():
    from quan_ultralytics_b200.optim import ClipSGD
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ClipSGD([{"params": [torch.zeros(3, requires_grad=True)], "lr": 0.1}])


def test_chunk_struct_matches_the_header():
    from quan_ultralytics_b200 import optim
    assert optim._CHUNK_DTYPE.itemsize == 32 and optim._CHUNK_DTYPE.fields["off"][1] == 16 and optim._CHUNK_DTYPE.fields["grp"][1] == 28
