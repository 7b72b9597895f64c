"""SURVEY.md §8 row a12 on cpu AND cuda, forward and backward: conv1d / max_pool1d / avg_pool1d against what the unmodified reference
computes (tests/golden/family_1d.npz, tests/golden/make_golden_1d.py). conv1d cases have n_out == kernel size — the only inputs for
which the reference's expression is defined (elsewhere it raises; this package then computes the ordinary strided correlation, checked
here against an explicit NumPy loop). The pooling functions return (N, C, k) like the reference (it reduces the window POSITIONS)."""
import os

import numpy as np
import pytest

import pydynet_b200 as pdn
import pydynet_b200.nn.functional as F

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "family_1d.npz"))
DEVICES = ["cpu", pytest.param("cuda:0", marks=pytest.mark.gpu)]
f32 = np.float32


def T(a, dev, rg=False):
    return pdn.Tensor(np.asarray(a), dtype=np.asarray(a).dtype, device=dev, requires_grad=rg)


def host(a):
    return np.asarray(a.get() if hasattr(a, "get") else a)


def close(got, ref, what):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    assert got.shape == ref.shape, f"{what}: {got.shape} vs {ref.shape}"
    err = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-30)
    assert err < 1e-4, f"{what}: normwise rel err {err:.3e}"  # north_star tolerance for fp32


@pytest.mark.parametrize("dev", DEVICES)
@pytest.mark.parametrize("i", [0, 1, 2])
def test_conv1d_matches_the_reference_where_it_is_defined(dev, i):
    stride, pad = (int(v) for v in G[f"conv1d{i}.cfg"])
    x, w = T(G[f"conv1d{i}.x"], dev, True), T(G[f"conv1d{i}.w"], dev, True)
    out = F.conv1d(x, w, pad, stride)
    close(out.numpy(), G[f"conv1d{i}.out"], "out")
    (out * T(G[f"conv1d{i}.g"], dev)).sum().backward()
    close(host(x.grad), G[f"conv1d{i}.dx"], "dx")
    close(host(w.grad), G[f"conv1d{i}.dw"], "dw")


@pytest.mark.parametrize("dev", DEVICES)
def test_conv1d_general_shape_is_the_strided_correlation(dev):
    rng = np.random.default_rng(3)
    x, w = rng.standard_normal((2, 3, 17)).astype(f32), rng.standard_normal((5, 3, 4)).astype(f32)
    stride, pad = 2, 1
    xp = np.pad(x, ((0, 0), (0, 0), (pad, pad)))
    L = (17 + 2 * pad - 4) // stride + 1
    ref = np.zeros((2, 5, L))
    for l in range(L):
        ref[:, :, l] = np.einsum("ncj,ocj->no", xp[:, :, l * stride:l * stride + 4].astype(np.float64), w.astype(np.float64))
    tx, tw = T(x, dev, True), T(w, dev, True)
    out = F.conv1d(tx, tw, pad, stride)
    close(out.numpy(), ref, "out")
    out.sum().backward()
    assert host(tx.grad).shape == x.shape and host(tw.grad).shape == w.shape


@pytest.mark.parametrize("dev", DEVICES)
@pytest.mark.parametrize("i", [0, 1, 2])
@pytest.mark.parametrize("mode", ["max", "avg"])
def test_pool1d_forward_backward(dev, i, mode):
    k, stride, pad = (int(v) for v in G[f"pool1d{i}.cfg"])
    x = T(G[f"pool1d{i}.x"], dev, True)
    out = (F.max_pool1d if mode == "max" else F.avg_pool1d)(x, k, stride, pad)
    close(out.numpy(), G[f"pool1d{i}.{mode}.out"], "out")
    (out * T(G[f"pool1d{i}.{mode}.g"], dev)).sum().backward()
    close(host(x.grad), G[f"pool1d{i}.{mode}.dx"], "dx")
