"""ClipSGD / ParamEMA (quan_sgd_clip_step, quan_ema_update) against the numpy oracle and against the torch calls the reference's
trainer makes (clip_grad_norm_ + torch.optim.SGD(nesterov) — engine/trainer.py:586-594, :799-806), over several steps, with
clipping active and inactive, parameters without gradient, odd sizes (scalar tails, unaligned chunks) and multi-chunk tensors."""
import numpy as np
import pytest
import torch

from oracle import quan_oracle as O

pytestmark = pytest.mark.gpu


def _make(seed, dev="cuda"):
    g = torch.Generator().manual_seed(seed)
    shapes = [(33, 7, 3, 3), (5,), (1,), (64, 4), (20000,), (3, 3), (8193,), (16, 16, 1, 1)]
    return [torch.randn(s, generator=g).to(dev).requires_grad_(True) for s in shapes]


@pytest.mark.parametrize("max_norm,scale", [(10.0, 0.01), (10.0, 5.0), (0.0, 1.0)])
@pytest.mark.parametrize("nesterov", [True, False])
def test_clip_sgd_matches_torch_and_oracle(max_norm, scale, nesterov):
    from quan_ultralytics_b200.optim import ClipSGD
    ours, ref = _make(0), _make(0)
    grp = [0, 1, 2, 0, 1, 2, 0, 1]
    lrs, wds = [0.01, 0.02, 0.005], [0.0, 5e-4, 1e-2]
    mk = lambda ps: [{"params": [p for p, g in zip(ps, grp) if g == gi], "lr": lrs[gi], "weight_decay": wds[gi]} for gi in range(3)]
    opt = ClipSGD(mk(ours), momentum=0.937, nesterov=nesterov, max_norm=max_norm)
    topt = torch.optim.SGD(mk(ref), lr=0.1, momentum=0.937, nesterov=nesterov)
    P = [p.detach().double().cpu().numpy() for p in ours]
    Bf = [np.zeros_like(p) for p in P]
    gen = torch.Generator().manual_seed(1)
    for step in range(4):
        grads = [torch.randn(p.shape, generator=gen) * scale for p in ours]
        skip = 2 if step % 2 else None                                   # a parameter without gradient this step
        for i, (po, pr, g) in enumerate(zip(ours, ref, grads)):
            po.grad = None if i == skip else g.to("cuda").clone()
            pr.grad = None if i == skip else g.to("cuda").clone()
        if max_norm > 0:
            tn = torch.nn.utils.clip_grad_norm_(ref, max_norm)
        topt.step()
        opt.step()
        torch.cuda.synchronize()
        idx = [i for i in range(len(P)) if i != skip]
        np_, nb_, ng_, total = O.sgd_clip_step([P[i] for i in idx], [grads[i].double().numpy() for i in idx], [Bf[i] for i in idx],
                                               [grp[i] for i in idx], lrs, wds, 0.937, max_norm, nesterov)
        for j, i in enumerate(idx):
            P[i], Bf[i] = np_[j], nb_[j]
        if max_norm > 0:
            assert abs(float(opt.total_norm) - total) <= 1e-5 * total and abs(float(tn) - total) <= 1e-5 * total
        for i, (po, pr) in enumerate(zip(ours, ref)):
            np.testing.assert_allclose(po.detach().cpu().numpy(), P[i], rtol=2e-5, atol=2e-6)
            torch.testing.assert_close(po.detach(), pr.detach(), rtol=2e-5, atol=2e-6)
            if i != skip and not nesterov:      # (torch's foreach SGD adds momentum*buf INTO .grad when nesterov: not comparable)
                torch.testing.assert_close(po.grad, pr.grad, rtol=1e-5, atol=1e-7)        # clip_grad_norm_ scales .grad in place
                torch.testing.assert_close(opt.momentum_of(po), topt.state[pr]["momentum_buffer"], rtol=2e-5, atol=2e-6)


def test_clip_sgd_zero_grad_and_graph_capture():
    from quan_ultralytics_b200.optim import ClipSGD
    ps, ref = _make(3), _make(3)
    for p, r in zip(ps, ref):
        p.grad = torch.randn_like(p)
        r.grad = p.grad.clone()
    opt = ClipSGD([{"params": ps, "lr": 0.05, "weight_decay": 1e-3}], momentum=0.9, max_norm=1.0)
    topt = torch.optim.SGD([{"params": ref, "lr": 0.05, "weight_decay": 1e-3}], momentum=0.9, nesterov=True)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            opt.step()
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):                                 # replays: the same gradients applied three times
        g.replay()
        torch.nn.utils.clip_grad_norm_(ref, 1.0)       # idempotent once clipped (norm == max_norm -> coefficient ~1)
        topt.step()
    torch.cuda.synchronize()
    # both sides leave the gradients scaled in place, so the repeated clip compounds identically
    for p, r in zip(ps, ref):
        torch.testing.assert_close(p.detach(), r.detach(), rtol=1e-4, atol=1e-5)
    opt.step(zero_grad=True)
    torch.cuda.synchronize()
    assert all(float(p.grad.abs().max()) == 0.0 for p in ps)


def test_param_ema_matches_reference_rule():
    from quan_ultralytics_b200.optim import ParamEMA
    net = torch.nn.Sequential(torch.nn.Linear(37, 19), torch.nn.BatchNorm1d(19)).cuda()
    ema = ParamEMA(net, decay=0.9999, tau=2000.0)
    want = {k: v.detach().double().cpu().numpy().copy() for k, v in net.state_dict().items() if v.dtype == torch.float32}
    for it in range(1, 4):
        with torch.no_grad():
            for p in net.parameters():
                p.add_(torch.randn_like(p) * 0.1)
            net[1].running_mean.add_(0.3)
        ema.update()
        d = 0.9999 * (1 - np.exp(-it / 2000.0))
        for k in want:
            want[k] = O.ema_update(want[k], net.state_dict()[k].detach().double().cpu().numpy(), d)
    torch.cuda.synchronize()
    for k, v in ema.state_dict().items():
        np.testing.assert_allclose(v.cpu().numpy(), want[k], rtol=1e-5, atol=1e-6)
