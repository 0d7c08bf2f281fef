"""N > 1 host logic on CPU (gloo, world_size 2): the synced-IQBN protocol in functional._IQBNTrain — partial sums ->
all_reduce -> finalize; local parameter gradients, global dx — checked against the single-process oracle on the
concatenated batch.  The CUDA kernels are replaced by oracle-backed CPU emulations of the SAME ops-level interface
(monkeypatched in the test only), so what is exercised is the product's rank logic, counts and collectives."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import quan_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _emulated_ops():
    """CPU emulation of the ops.iqbn_* entry points (numpy oracle inside), same signatures as quan_ultralytics_b200.ops."""
    from quan_ultralytics_b200 import ops

    def n(t):
        return t.detach().double().numpy()

    def partial_sums(x, layout):
        a = n(x)
        s = a.sum(axis=(0, 2, 3)).reshape(-1)
        ss = (a * a).sum(axis=(0, 2, 3)).reshape(-1)
        return torch.from_numpy(np.concatenate([s, ss]))

    def finalize_stats(sums, count, C_, gamma, beta, eps, momentum, rm, rv):
        s = sums.numpy()
        mean = s[:4 * C_] / count
        var = s[4 * C_:8 * C_] / count - mean * mean + 1e-8
        rstd = 1.0 / np.sqrt(var + eps)
        if rm is not None:
            rm.copy_(torch.from_numpy(((1 - momentum) * n(rm).reshape(-1) + momentum * mean).reshape(C_, 4)))
            rv.copy_(torch.from_numpy(((1 - momentum) * n(rv).reshape(-1) + momentum * var).reshape(C_, 4)))
        return torch.from_numpy(np.concatenate([mean, var, rstd, np.zeros(8 * C_)])).float()

    def train_stats(x, layout, gamma, beta, eps, momentum, rm, rv):
        B, C_, H, W, _ = x.shape
        return finalize_stats(partial_sums(x, layout), float(B * H * W), C_, gamma, beta, eps, momentum, rm, rv)

    def _unpack(stats, C_):
        s = stats.double().numpy()
        return s[:4 * C_].reshape(C_, 4), s[8 * C_:12 * C_].reshape(C_, 4)

    def apply_fwd(x, layout, stats, gamma, beta, act):
        C_ = x.size(1)
        mean, rstd = _unpack(stats, C_)
        bc = lambda a: a[None, :, None, None, :]
        z = (n(x) - bc(mean)) * bc(rstd) * bc(n(gamma)) + bc(n(beta))
        return torch.from_numpy(O.silu(z) if act else z).to(x.dtype)

    def bwd_reduce(dy, x, layout, stats, gamma, beta, act, count=0.0):
        C_ = x.size(1)
        mean, rstd = _unpack(stats, C_)
        bc = lambda a: a[None, :, None, None, :]
        xhat = (n(x) - bc(mean)) * bc(rstd)
        dz = n(dy) * (O.silu_grad(xhat * bc(n(gamma)) + bc(n(beta))) if act else 1.0)
        out = np.zeros(14 * C_)
        out[:4 * C_] = dz.sum(axis=(0, 2, 3)).reshape(-1)
        out[4 * C_:8 * C_] = (dz * xhat).sum(axis=(0, 2, 3)).reshape(-1)
        return torch.from_numpy(out)

    def bwd_coef(sums, count, stats, gamma):
        return None

    def bwd_apply(dy, x, layout, stats, gamma, beta, act, sums, count, want_param_grads=True, mix_t=None):
        C_ = x.size(1)
        mean, rstd = _unpack(stats, C_)
        bc = lambda a: a[None, :, None, None, :]
        xhat = (n(x) - bc(mean)) * bc(rstd)
        dz = n(dy) * (O.silu_grad(xhat * bc(n(gamma)) + bc(n(beta))) if act else 1.0)
        s = sums.numpy()
        sdz, sdzx = s[:4 * C_].reshape(C_, 4), s[4 * C_:8 * C_].reshape(C_, 4)
        dx = bc(n(gamma) * rstd) * (dz - bc(sdz) / count - xhat * bc(sdzx) / count)
        dg = torch.from_numpy(sdzx).float() if want_param_grads else None
        db = torch.from_numpy(sdz).float() if want_param_grads else None
        return torch.from_numpy(dx).to(x.dtype), dg, db

    return dict(as_layout=lambda x, layout=None: (x.contiguous(), ops.LAYOUT_BCHWQ), _f32c=ops._f32c,
                iqbn_partial_sums=partial_sums, iqbn_finalize_stats=finalize_stats, iqbn_train_stats=train_stats,
                iqbn_apply_fwd=apply_fwd, iqbn_bwd_reduce=bwd_reduce, iqbn_bwd_coef=bwd_coef, iqbn_bwd_apply=bwd_apply)


def _worker(rank, world, port, x_all, dy_all, gamma, beta, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import quan_ultralytics_b200 as Q
        from quan_ultralytics_b200 import functional as QF
        from quan_ultralytics_b200.distributed import convert_sync_iqbn
        for k, v in _emulated_ops().items():
            setattr(QF.ops, k, v)
        per = x_all.shape[0] // world
        x = torch.from_numpy(x_all[rank * per:(rank + 1) * per]).requires_grad_(True)
        dy = torch.from_numpy(dy_all[rank * per:(rank + 1) * per])
        bn = Q.IQBN(x.shape[1] * 4).double()
        with torch.no_grad():
            bn.gamma.copy_(torch.from_numpy(gamma))
            bn.beta.copy_(torch.from_numpy(beta))
        convert_sync_iqbn(bn)
        assert bn.sync and bn.process_group is None
        bn.train()
        y = bn(x, Q.ACT_SILU)
        y.backward(dy)
        # DDP would average the LOCAL parameter gradients: emulate that reduction here
        gg, gb = bn.gamma.grad.clone(), bn.beta.grad.clone()
        dist.all_reduce(gg)
        dist.all_reduce(gb)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), y=y.detach().numpy(), dx=x.grad.numpy(), dgamma=gg.numpy(),
                 dbeta=gb.numpy(), rm=bn.running_mean.numpy(), rv=bn.running_var.numpy(),
                 nbt=int(bn.num_batches_tracked))
    finally:
        dist.destroy_process_group()


def test_synced_iqbn_matches_single_process_on_the_global_batch(tmp_path):
    world = 2
    rng = np.random.default_rng(0)
    x = rng.normal(size=(4, 3, 5, 4, 4)) * 1.5 + 0.3
    x[2:] += 0.8                                  # make the two ranks' local statistics differ
    dy = rng.normal(size=x.shape)
    gamma, beta = rng.normal(size=(3, 4)) * 0.3 + 1, rng.normal(size=(3, 4)) * 0.2
    mp.spawn(_worker, args=(world, _free_port(), x, dy, gamma, beta, str(tmp_path)), nprocs=world, join=True)
    y_ref, rm, rv, _ = O.iqbn_train_fwd(x, gamma, beta, np.zeros((3, 4)), np.ones((3, 4)), act=True)
    dx_ref, dg_ref, db_ref = O.iqbn_train_bwd(dy, x, gamma, beta, act=True)
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        sl = slice(r * 2, r * 2 + 2)
        np.testing.assert_allclose(d["y"], y_ref[sl], rtol=1e-5, atol=1e-6)        # stats travel as float32, as in the product
        np.testing.assert_allclose(d["dx"], dx_ref[sl], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(d["dgamma"], dg_ref, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(d["dbeta"], db_ref, rtol=1e-5, atol=1e-5)
        np.testing.assert_allclose(d["rm"], rm, rtol=1e-6, atol=1e-7)              # every rank holds the GLOBAL stats
        np.testing.assert_allclose(d["rv"], rv, rtol=1e-6, atol=1e-7)
        assert int(d["nbt"]) == 1


def test_convert_sync_iqbn_marks_only_iqbn():
    import quan_ultralytics_b200 as Q
    from quan_ultralytics_b200.distributed import convert_sync_iqbn
    net = torch.nn.Sequential(Q.Conv(16, 16, 3, 1), Q.QUpsample(2), Q.Conv(16, 32, 1, 1))
    assert not any(m.sync for m in net.modules() if isinstance(m, Q.IQBN))
    convert_sync_iqbn(net, process_group="grp")
    bns = [m for m in net.modules() if isinstance(m, Q.IQBN)]
    assert len(bns) == 2 and all(m.sync and m.process_group == "grp" for m in bns)
