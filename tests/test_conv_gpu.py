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
