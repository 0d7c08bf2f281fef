"""Pin the CPU oracle (oracle/quan_oracle.py, oracle/torch_port.py) against outputs of the REAL reference
(tests/golden/quan_layers.npz, produced by tests/golden/make_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from oracle import quan_oracle as O
from oracle import torch_port as TP

CONV_CASES = ["k3s1", "k3s2", "k1", "dw", "g2d2", "k7s2"]
TOL = dict(rtol=1e-10, atol=1e-10)


def _conv_rec(golden, name):
    g = lambda k: golden[f"{name}/{k}"]
    cin, cout, k, s, p, d, grp, bias = [int(v) for v in g("conf")]
    w = [g("w_r"), g("w_i"), g("w_j"), g("w_k")]
    b = g("bias_r") if bias else None
    return g, w, b, (s, p, d, grp)


@pytest.mark.parametrize("mix", ["A", "B"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_qconv_fwd_bwd_matches_reference(golden, mix, case):
    g, w, b, (s, p, d, grp) = _conv_rec(golden, f"conv{mix}_{case}")
    y = O.qconv2d_fwd(g("x"), w, b, s, p, d, grp, O.MIX[mix])
    np.testing.assert_allclose(y, g("y"), **TOL)
    dx, dws, db = O.qconv2d_bwd(g("dy"), g("x"), w, s, p, d, grp, O.MIX[mix], has_bias=b is not None)
    np.testing.assert_allclose(dx, g("dx"), **TOL)
    for q, n in enumerate("rijk"):
        np.testing.assert_allclose(dws[q], g(f"dw_{n}"), **TOL)
    if b is not None:
        np.testing.assert_allclose(db, g("db_r"), **TOL)


@pytest.mark.parametrize("mix", ["A", "B"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_torch_port_matches_reference(golden, mix, case):
    g, w, b, (s, p, d, grp) = _conv_rec(golden, f"conv{mix}_{case}")
    t = lambda a: None if a is None else torch.from_numpy(a)
    y = TP.qconv2d(t(g("x")), *[t(v) for v in w], t(b), (s, s), (p, p), (d, d), grp, mix)
    np.testing.assert_allclose(y.numpy(), g("y"), **TOL)


def test_mix_matrices_differ_as_surveyed():
    # M_B is Hadamard (M^T M = 4I), M_A is not a sign-flip of it.
    assert np.allclose(O.M_B.T @ O.M_B, 4 * np.eye(4))
    assert not np.allclose(O.M_A.T @ O.M_A, 4 * np.eye(4))


def test_poincare(golden):
    q = O.poincare_fwd(golden["poincare/rgb"])
    np.testing.assert_allclose(q, golden["poincare/q"], **TOL)
    np.testing.assert_allclose(np.linalg.norm(q, axis=-1), 1.0, atol=1e-12)       # unit quaternions
    np.testing.assert_allclose(O.poincare_bwd(golden["poincare/rgb"], golden["poincare/gq"]),
                               golden["poincare/grgb"], **TOL)
    np.testing.assert_allclose(O.poincare_fwd(golden["poincare_n/rgb"]), golden["poincare_n/q"], **TOL)
    np.testing.assert_allclose(TP.poincare(torch.from_numpy(golden["poincare/rgb"])).numpy(), golden["poincare/q"],
                               **TOL)
    w = [golden[f"poincare/w_{n}"] for n in "rijk"]
    np.testing.assert_allclose(O.qconv2d_fwd(q, w, None, 2, 1, 1, 1, O.M_A), golden["poincare/y"], **TOL)


@pytest.mark.parametrize("tag", ["iqbnA", "iqbnB"])
def test_iqbn(golden, tag):
    g = lambda k: golden[f"{tag}/{k}"]
    y, rm, rv, _ = O.iqbn_train_fwd(g("x"), g("gamma"), g("beta"), g("rm0"), g("rv0"))
    np.testing.assert_allclose(y, g("y"), **TOL)
    np.testing.assert_allclose(rm, g("rm1"), **TOL)
    np.testing.assert_allclose(rv, g("rv1"), **TOL)
    dx, dgam, dbet = O.iqbn_train_bwd(g("dy"), g("x"), g("gamma"), g("beta"))
    np.testing.assert_allclose(dx, g("dx"), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(dgam, g("dgamma"), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(dbet, g("dbeta"), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(O.iqbn_eval_fwd(g("x"), g("gamma"), g("beta"), g("rm1"), g("rv1")), g("y_eval"), **TOL)
    assert int(g("nbt")) == 1


def test_conv_block(golden):
    g = lambda k: golden[f"block/{k}"]
    w = [g(f"w_{n}") for n in "rijk"]
    c = O.qconv2d_fwd(g("x"), w, None, 1, 1, 1, 1, O.M_A)
    C = c.shape[1]
    gamma, beta = np.ones((C, 4)), np.zeros((C, 4))
    y, rm, rv, _ = O.iqbn_train_fwd(c, gamma, beta, np.zeros((C, 4)), np.ones((C, 4)), act=True)
    np.testing.assert_allclose(y, g("y"), rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(rm, g("rm1"), **TOL)
    np.testing.assert_allclose(rv, g("rv1"), **TOL)
    dc, dgam, dbet = O.iqbn_train_bwd(g("dy"), c, gamma, beta, act=True)
    np.testing.assert_allclose(dgam, g("dgamma"), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(dbet, g("dbeta"), rtol=1e-8, atol=1e-9)
    dx, dws, _ = O.qconv2d_bwd(dc, g("x"), w, 1, 1, 1, 1, O.M_A)
    np.testing.assert_allclose(dx, g("dx"), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(dws[0], g("dw_r"), rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(dws[3], g("dw_k"), rtol=1e-8, atol=1e-9)
    # torch port of the whole block
    blk = TP.Conv(4, 8, 3, 1, mix="A").double()
    with torch.no_grad():
        for n in "rijk":
            getattr(blk.conv, f"weight_{n}").copy_(torch.from_numpy(g(f"w_{n}")))
    np.testing.assert_allclose(blk(torch.from_numpy(g("x"))).detach().numpy(), g("y"), rtol=1e-9, atol=1e-10)


def test_upsample(golden):
    g = lambda k: golden[f"upsample/{k}"]
    np.testing.assert_allclose(O.qupsample_fwd(g("x"), 2), g("y"), **TOL)
    np.testing.assert_allclose(O.qupsample_bwd(g("dy"), 2), g("dx"), **TOL)
    np.testing.assert_allclose(TP.qupsample(torch.from_numpy(g("x")), 2).numpy(), g("y"), **TOL)


def test_eval_bwd_matches_autograd():
    rng = np.random.default_rng(0)
    x, dy = rng.normal(size=(2, 3, 4, 4, 4)), rng.normal(size=(2, 3, 4, 4, 4))
    gam, bet, rm, rv = rng.normal(size=(3, 4)), rng.normal(size=(3, 4)), rng.normal(size=(3, 4)), rng.random((3, 4)) + .5
    xt = torch.from_numpy(x).requires_grad_(True)
    v = lambda a: torch.from_numpy(a).view(1, 3, 1, 1, 4)
    z = (xt - v(rm)) / torch.sqrt(v(rv) + 1e-5) * v(gam) + v(bet)
    torch.nn.functional.silu(z).backward(torch.from_numpy(dy))
    np.testing.assert_allclose(O.iqbn_eval_bwd(dy, x, gam, bet, rm, rv, act=True), xt.grad.numpy(), rtol=1e-9, atol=1e-10)


def test_oracle_matches_reference_qwrn_trace():
    """BASELINE config[0]: layers recorded from the reference's Q-WRN-16-2 training step (make_qwrn_trace.py) through the
    numpy oracle (classification flavour: mixing matrix M_B, bias on S_r)."""
    from pathlib import Path
    t = np.load(Path(__file__).resolve().parent / "golden" / "qwrn_trace.npz")
    for tag in ("conv_s2", "conv_deep"):
        g = lambda k: t[f"{tag}/{k}"].astype(np.float64)
        k, s, p, d, grp, has_bias = [int(v) for v in t[f"{tag}/conf"]]
        w = [g("w_r"), g("w_i"), g("w_j"), g("w_k")]
        y = O.qconv2d_fwd(g("x"), w, g("bias_r") if has_bias else None, s, p, d, grp, O.M_B)
        assert np.max(np.abs(y - g("y"))) / np.max(np.abs(g("y"))) <= 1e-5      # fixtures are stored in fp32
        dx, dw, db = O.qconv2d_bwd(g("dy"), g("x"), w, s, p, d, grp, O.M_B, has_bias=bool(has_bias))
        assert np.max(np.abs(dx - g("dx"))) / np.max(np.abs(g("dx"))) <= 1e-5
        assert np.max(np.abs(dw[2] - g("dw_j"))) / np.max(np.abs(g("dw_j"))) <= 1e-5
    for tag in ("bn_first", "bn_last"):
        g = lambda k: t[f"{tag}/{k}"].astype(np.float64)
        y, _, _, _ = O.iqbn_train_fwd(g("x"), g("gamma"), g("beta"), eps=float(t[f"{tag}/eps"]))
        assert np.max(np.abs(y - g("y"))) / np.max(np.abs(g("y"))) <= 1e-5
        dx, dg, db = O.iqbn_train_bwd(g("dy"), g("x"), g("gamma"), g("beta"), eps=float(t[f"{tag}/eps"]))
        assert np.max(np.abs(dx - g("dx"))) / np.max(np.abs(g("dx"))) <= 2e-4
        assert np.max(np.abs(dg - g("dgamma"))) / np.max(np.abs(g("dgamma"))) <= 2e-4
