"""pydynet_b200.cuda.graphed_step on the GPU: a training step recorded into a CUDA graph and replayed must walk the SAME trajectory
as the eager step (same kernels, same order; Adam's step counter and bias correction live in device memory during replays) — LeNet
(convolutions, pooling, fused cross-entropy, flat Adam) over 6 steps with fresh inputs every step, and the config-1 matmul
forward + backward."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lenet_run(graphed, steps=6):
    import pydynet_b200 as pdn
    from pydynet_b200.optim import Adam
    from workloads.lenet import ConvNet, train_step
    dev = "cuda:0"
    np.random.seed(42)
    net = ConvNet().to(dev)
    opt = Adam(net.parameters(), lr=1e-3)
    net.train()
    rng = np.random.default_rng(0)
    batches = [(rng.random((32, 1, 28, 28)).astype(np.float32), rng.integers(0, 10, 32)) for _ in range(steps)]
    fn = lambda X, y: train_step(net, opt, X, y)
    step = pdn.cuda.graphed_step(fn, optimizers=[opt]) if graphed else fn
    losses = []
    for i, (X, y) in enumerate(batches):
        if i == 4:
            opt.lr = 5e-4  # a scheduler changing the learning rate between steps must reach the recorded update
        loss = step(pdn.Tensor(X, dtype=np.float32, device=dev), pdn.Tensor(y, device=dev))
        losses.append(float(loss.item()))
    return losses, {k: p.numpy() for k, p in net._parameters.items()}, opt.t


def test_lenet_graphed_equals_eager():
    le, pe, te = _lenet_run(False)
    lg, pg, tg = _lenet_run(True)
    assert te == tg == 7
    np.testing.assert_allclose(lg, le, rtol=1e-5)
    for k in pe:
        err = np.linalg.norm(pg[k] - pe[k]) / max(np.linalg.norm(pe[k]), 1e-30)
        assert err < 1e-5, (k, err)


def test_matmul_fwd_bwd_graphed():
    import pydynet_b200 as pdn
    dev = "cuda:0"
    rng = np.random.default_rng(0)
    A, B = rng.standard_normal((512, 512)).astype(np.float32), rng.standard_normal((512, 512)).astype(np.float32)
    x, w = pdn.Tensor(A, dtype=A.dtype, device=dev, requires_grad=True), pdn.Tensor(B, dtype=B.dtype, device=dev, requires_grad=True)

    def fn(xx, ww):
        x.zero_grad()
        w.zero_grad()
        out = pdn.matmul(x, w)
        out.sum().backward()
        return out

    step = pdn.cuda.graphed_step(fn)
    for _ in range(5):
        out = step(x, w)
    ones = np.ones((512, 512))
    ref = A.astype(np.float64) @ B
    assert np.linalg.norm(out.numpy() - ref) / np.linalg.norm(ref) < 1e-4
    gx, gw = ones @ B.astype(np.float64).T, A.astype(np.float64).T @ ones
    assert np.linalg.norm(x.grad.get() - gx) / np.linalg.norm(gx) < 1e-4
    assert np.linalg.norm(w.grad.get() - gw) / np.linalg.norm(gw) < 1e-4
