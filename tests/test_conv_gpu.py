"""GPU parity of Conv2d forward / backward-data / backward-weight against the oracle's im2col restatement of the reference
(functional.py:254-281) in float64. Covers the TMA-tiled stride-1 path (csrc/conv_tma.cu: >= 16 contraction channels) with ragged
sizes — widths that are not multiples of the 16-pixel box, channel counts that are not multiples of the 64-channel block or the
16-wide MMA k-step, kernel sizes 1/3/5, paddings 0..2 — and the gather path it falls back to (stride 2, few channels).
Tolerance 1e-4 normwise (BASELINE north_star)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [  # N, C, H, W, O, k, stride, pad
    (4, 20, 14, 14, 50, 3, 1, 1), (2, 64, 56, 56, 128, 3, 1, 1), (3, 33, 19, 23, 70, 3, 1, 0), (2, 16, 9, 40, 24, 5, 1, 2),
    (5, 96, 7, 7, 200, 1, 1, 0), (2, 130, 12, 17, 300, 3, 1, 1), (3, 24, 30, 30, 40, 3, 2, 1), (4, 3, 20, 20, 32, 3, 1, 1),
    (2, 18, 11, 13, 10, 3, 1, 2),
]


def _err(got, ref):
    return float(np.linalg.norm(got.astype(np.float64) - ref) / max(np.linalg.norm(ref), 1e-30))


@pytest.mark.parametrize("case", CASES)
def test_conv2d_fwd_bwd(case):
    import pydynet_b200 as pdn
    import pydynet_b200.nn.functional as F
    from oracle import pdn_oracle as O_
    N, C, H, W, O, k, stride, pad = case
    rng = np.random.default_rng(sum(case))
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((O, C, k, k)) / np.sqrt(C * k * k)).astype(np.float32)
    b = rng.standard_normal((1, O, 1, 1)).astype(np.float32)
    oh, ow = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    g = rng.standard_normal((N, O, oh, ow)).astype(np.float32)
    ref = O_.conv2d_fwd_bwd(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), g.astype(np.float64), stride, pad)
    dev = "cuda:0"
    tx, tw, tb = (pdn.Tensor(t, dtype=np.float32, device=dev, requires_grad=True) for t in (x, w, b))
    out = F.conv2d(tx, tw, pad, stride) + tb
    (out * pdn.Tensor(g, dtype=np.float32, device=dev)).sum().backward()
    got = (out.numpy(), tx.grad.get(), tw.grad.get(), tb.grad.get())
    for name, a, r in zip(("out", "dx", "dw", "db"), got, ref):
        assert a.shape == r.shape, name
        assert np.isfinite(a).all(), name
        e = _err(a, r)
        assert e < 1e-4, f"conv {case} {name}: normwise rel err {e:.3e}"


@pytest.mark.parametrize("shape", [(5, 1, 28, 28, 20, 3, 1), (3, 3, 17, 13, 7, 3, 0), (4, 1, 12, 9, 6, 5, 2), (2, 2, 8, 8, 64, 3, 2),
                                   (300, 1, 28, 28, 20, 3, 1)])
def test_thin_first_layer_kernels(shape):
    """Direct fp32 kernels for stride-1 convolutions over 1-3 input channels (k_conv_thin_fwd / k_conv_thin_bwd_weight: the first
    layer of BASELINE config 2 is 1 -> 20 channels, 3x3): output, weight gradient and bias gradient against an fp64 evaluation of
    functional.py:254-281, 5e-6 normwise (plain fp32 FMA chains, no BF16 split); the launch counter proves the direct kernels ran."""
    import pydynet_b200 as pdn
    import pydynet_b200.nn as nn
    from pydynet_b200.backend import lib
    N, C, H, W, Oc, k, pad = shape
    rng = np.random.default_rng(0)
    x = rng.standard_normal((N, C, H, W)).astype(np.float32)
    w = (rng.standard_normal((Oc, C, k, k)) * .3).astype(np.float32)
    b = rng.standard_normal(Oc).astype(np.float32)
    conv = nn.Conv2d(C, Oc, k, 1, pad, dtype=np.float32).to("cuda:0")
    with conv.weight.device:
        conv.weight.data[...] = w
        conv.bias.data[...] = b.reshape(1, -1, 1, 1)
    tx = pdn.Tensor(x, dtype=np.float32, device="cuda:0")
    lib.watch_launches("conv_thin_fwd")
    y = conv(tx)
    assert lib.watched_launch_count() == 1
    gyv = rng.standard_normal(y.shape).astype(np.float32)
    lib.watch_launches("conv_thin_bwd_weight")
    (y * pdn.Tensor(gyv, dtype=np.float32, device="cuda:0")).sum().backward()
    assert lib.watched_launch_count() == 1
    lib.watch_launches(None)
    xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (pad, pad), (pad, pad)))
    oh, ow = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    yr, dwr = np.zeros((N, Oc, oh, ow)), np.zeros((Oc, C, k, k))
    for ky in range(k):
        for kx in range(k):
            patch = xp[:, :, ky:ky + oh, kx:kx + ow]
            yr += np.einsum("nchw,oc->nohw", patch, w[:, :, ky, kx].astype(np.float64))
            dwr[:, :, ky, kx] = np.einsum("nohw,nchw->oc", gyv.astype(np.float64), patch)
    yr += b[None, :, None, None]

    def err(a, r):
        a = np.asarray(a.get() if hasattr(a, "get") else a, np.float64).reshape(r.shape)
        return np.linalg.norm(a - r) / np.linalg.norm(r)

    assert err(y.numpy(), yr) < 5e-6
    assert err(conv.weight.grad, dwr) < 5e-6
    assert err(conv.bias.grad, gyv.sum((0, 2, 3))) < 5e-6
