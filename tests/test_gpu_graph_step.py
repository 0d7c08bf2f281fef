"""GraphedTrainStep (forward graph -> eager reference loss -> backward + ClipSGD graph) against the same step run eagerly, on the
real QUAN-YOLO11n-OBB graph: same losses and same parameters after several optimizer steps."""
import pytest
import torch

from quan_ultralytics_b200 import refenv

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(refenv.find_reference() is None, reason="no reference tree (baseline/_ref)")]


def _build(seed):
    from quan_ultralytics_b200 import optim, workloads
    torch.manual_seed(seed)
    model = workloads.build_yolo_obb("n", 15, "cuda", swapped=True).train()
    return model, optim.yolo_clip_sgd(model)


@pytest.mark.parametrize("autocast", [None, torch.bfloat16])
def test_graphed_yolo_step_equals_eager(autocast):
    from quan_ultralytics_b200 import install as qi
    from quan_ultralytics_b200 import workloads
    from quan_ultralytics_b200.graphs import GraphedTrainStep
    import quan_ultralytics_b200 as Q
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        m_e, o_e = _build(0)
        m_g, o_g = _build(0)
        m_g.load_state_dict(m_e.state_dict())
        if autocast is None:
            # exact arithmetic on both sides (fp32 CUDA-core engine, TF32 off for the QER convolutions): the only difference left between
            # the eager and the captured step is the order of atomic split-K sums, so trajectories stay together
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
            for m in list(m_e.modules()) + list(m_g.modules()):
                if isinstance(m, Q.QConv2D):
                    m.algo = Q.ALGO_DIRECT
        batch = workloads.synthetic_obb_batch(2, 256, "cuda", boxes_per_image=8, seed=5)
        sd0 = {k: v.clone() for k, v in m_e.state_dict().items()}
        step = GraphedTrainStep(lambda img: m_g(img), lambda preds, img, b: m_g.loss(b, preds), o_g, [batch["img"]],
                                list(m_g.parameters()), autocast=autocast, loss_args=(batch,))
        m_g.load_state_dict(sd0)                       # warm-up moved the running statistics: start both from the same state
        losses_e, losses_g = [], []
        for it in range(3):
            with torch.autocast("cuda", dtype=autocast, enabled=autocast is not None):
                loss, _ = m_e(batch)
            o_e.zero_grad()
            loss.backward()
            o_e.step()
            losses_e.append(float(loss.detach()))
            lg, _ = step([batch["img"]], (batch,))
            losses_g.append(float(lg.detach()))
        torch.cuda.synchronize()
        # same kernels, same weights: the first loss agrees to rounding (split-K atomics order the wgrad / statistics sums differently
        # from run to run); after clipped SGD steps of norm 10 * lr the trajectories may drift by ~1e-4 in fp32
        tol0, tol = (1e-5, 1e-3) if autocast is None else (2e-2, 2e-2)
        assert abs(losses_e[0] - losses_g[0]) <= tol0 * abs(losses_e[0]), (losses_e, losses_g)
        for a, b in zip(losses_e, losses_g):
            assert abs(a - b) <= tol * abs(a), (losses_e, losses_g)
        if autocast is None:
            for (n, p), (_, q) in zip(m_e.named_parameters(), m_g.named_parameters()):
                assert float((p - q).abs().max()) <= 2e-3 * max(float(p.abs().max()), 1e-3), n
        assert losses_g[2] != losses_g[0]              # the captured optimizer really moves the weights
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
        qi.uninstall()
