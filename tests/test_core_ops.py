"""Black-box re-statement of the reference's own unit tests (reference tests/test_tensor_basic.py:87-117,
tests/test_ops_extended.py:15-115, tests/test_backward.py:20-73 — 77 cases, all NumPy-equality pins) against this
package, on BOTH devices: ``cpu`` (host engine logic, runs without a GPU) and ``cuda:0`` (libpdn_b200.so kernels,
marked gpu).  Tolerances are the reference's: np.allclose defaults / 1e-6 for fp64 unary ops."""
import random

import numpy as np
import pytest

import pydynet_b200 as pdn

DEVICES = [pytest.param("cpu", id="cpu"), pytest.param("cuda:0", id="cuda", marks=pytest.mark.gpu)]
TYPES = [np.float16, np.float32, np.float64]


def _np(t):
    return t.numpy() if isinstance(t, pdn.Tensor) else (t.get() if hasattr(t, "get") else np.asarray(t))


def _broadcast_pairs(n, seed):
    rnd, rng = random.Random(seed), np.random.default_rng(seed)
    for _ in range(n):
        nd = rnd.randint(0, 4)
        s1, s2 = [], []
        for _ in range(nd):
            if rnd.random() < 0.5:
                a, b = rnd.choice([(1, rnd.randint(1, 5)), (rnd.randint(1, 5), 1)])
            else:
                a = b = rnd.randint(1, 5)
            s1.append(a)
            s2.append(b)
        s1 = s1[rnd.randint(0, len(s1)):]
        yield (rng.standard_normal(s1).astype(rng.choice(TYPES)), rng.standard_normal(s2).astype(rng.choice(TYPES)))


def _matmul_pairs(n, seed):
    rnd, rng = random.Random(seed), np.random.default_rng(seed)
    for _ in range(n):
        nd = rnd.randint(0, 4)
        s1, s2 = [], []
        for _ in range(nd):
            if rnd.random() < 0.5:
                a, b = rnd.choice([(1, rnd.randint(1, 5)), (rnd.randint(1, 5), 1)])
            else:
                a = b = rnd.randint(1, 5)
            s1.append(a)
            s2.append(b)
        m, k, p = rnd.randint(1, 5), rnd.randint(1, 5), rnd.randint(1, 5)
        s1, s2 = s1 + [m, k], s2 + [k, p]
        s1 = s1[rnd.randint(0, len(s1) - 2):]
        yield (rng.standard_normal(s1).astype(rng.choice(TYPES)), rng.standard_normal(s2).astype(rng.choice(TYPES)))


BINARY = [("add", np.add), ("sub", np.subtract), ("mul", np.multiply), ("div", np.divide), ("pow", np.power),
          ("maximum", np.maximum), ("minimum", np.minimum)]


@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("name,np_func", BINARY)
@pytest.mark.filterwarnings("ignore:invalid value")
@pytest.mark.filterwarnings("ignore:divide by zero")
def test_binary_operator(device, name, np_func):
    for a, b in _broadcast_pairs(8, 42):
        out = getattr(pdn, name)(pdn.Tensor(a, device=device), pdn.Tensor(b, device=device))
        ref = np_func(a, b)
        assert out.shape == ref.shape
        assert out.dtype == ref.dtype
        assert np.allclose(_np(out), ref, equal_nan=True, rtol=2e-3 if ref.dtype == np.float16 else 1e-5)


@pytest.mark.parametrize("device", DEVICES)
def test_matmul_forward(device):
    for a, b in _matmul_pairs(12, 42):
        out = pdn.matmul(pdn.Tensor(a, device=device), pdn.Tensor(b, device=device))
        ref = np.matmul(a, b)
        assert out.shape == ref.shape
        assert out.dtype == ref.dtype
        assert np.allclose(_np(out), ref, equal_nan=True, rtol=1e-2 if ref.dtype == np.float16 else 1e-5, atol=1e-2 if ref.dtype == np.float16 else 1e-6)


UNARY = [("abs", np.abs), ("exp", np.exp), ("log", np.log), ("sign", np.sign), ("sigmoid", lambda x: 1 / (1 + np.exp(-x))),
         ("tanh", np.tanh), ("sqrt", np.sqrt), ("square", np.square)]


@pytest.mark.parametrize("device", DEVICES)
@pytest.mark.parametrize("name,np_func", UNARY)
def test_unary_forward(device, name, np_func):
    x = np.random.default_rng(123).uniform(0.1, 2.0, size=(3, 4)).astype(np.float64)
    if name in ("abs", "sign", "sigmoid", "tanh"):
        x = x - 1.0
    out = getattr(pdn, name)(pdn.Tensor(x, device=device))
    np.testing.assert_allclose(_np(out), np_func(x), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("device", DEVICES)
def test_reductions_forward(device):
    x = np.random.default_rng(123).standard_normal((2, 3, 4)).astype(np.float64)
    t = pdn.Tensor(x, device=device)
    for fn, ref, kw in [(pdn.sum, np.sum, {}), (pdn.sum, np.sum, {"axis": 1}), (pdn.sum, np.sum, {"axis": (0, 2), "keepdims": True}),
                        (pdn.mean, np.mean, {}), (pdn.mean, np.mean, {"axis": 2}), (pdn.mean, np.mean, {"axis": (0, 1), "keepdims": True}),
                        (pdn.max, np.max, {"axis": 1}), (pdn.min, np.min, {"axis": 2}), (pdn.max, np.max, {}),
                        (pdn.argmax, np.argmax, {"axis": 1}), (pdn.argmin, np.argmin, {"axis": 2})]:
        got = fn(t, **kw)
        exp = ref(x, **kw)
        assert got.shape == np.shape(exp)
        np.testing.assert_allclose(_np(got), exp, rtol=1e-12)
    assert pdn.argmax(t, axis=1).dtype == np.int64


@pytest.mark.parametrize("device", DEVICES)
def test_shape_ops(device):
    x = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    t = pdn.Tensor(x, device=device)
    np.testing.assert_array_equal(_np(pdn.reshape(t, (4, 6))), x.reshape(4, 6))
    np.testing.assert_array_equal(_np(t.reshape(6, 4)), x.reshape(6, 4))
    np.testing.assert_array_equal(_np(pdn.transpose(t, (1, 0, 2))), x.transpose(1, 0, 2))
    np.testing.assert_array_equal(_np(t.T), x.T)
    np.testing.assert_array_equal(_np(pdn.swapaxes(t, 0, 2)), x.swapaxes(0, 2))
    u = pdn.unsqueeze(t, (0, -1))
    assert u.shape == (1, 2, 3, 4, 1)
    assert pdn.squeeze(u, (0, 4)).shape == (2, 3, 4)
    assert pdn.squeeze(u).shape == (2, 3, 4)
    for axis in (0, 1, 2):
        n = x.shape[axis] if axis != 1 else 3
        parts = pdn.split(t, 2 if axis != 1 else 3, axis=axis)
        ref = np.split(x, 2 if axis != 1 else 3, axis=axis)
        assert len(parts) == len(ref)
        for p, r in zip(parts, ref):
            np.testing.assert_array_equal(_np(p), r)
        np.testing.assert_array_equal(_np(pdn.concat(parts, axis=axis)), x)
    parts = pdn.split(t, (1, 3), axis=2)
    for p, r in zip(parts, np.split(x, (1, 3), axis=2)):
        np.testing.assert_array_equal(_np(p), r)


@pytest.mark.parametrize("device", DEVICES)
def test_backward_scalar_polynomial(device):
    x = pdn.Tensor(2.0, device=device, requires_grad=True)
    y = x * x + 3 * x + 1
    y.backward()
    np.testing.assert_allclose(_np(x.grad), 7.0, rtol=1e-6)


@pytest.mark.parametrize("device", DEVICES)
def test_backward_broadcast_add(device):
    a = np.arange(6, dtype=np.float64).reshape(2, 3)
    b = np.array([1.0, 2.0, 3.0])
    x, y = pdn.Tensor(a, device=device, requires_grad=True), pdn.Tensor(b, device=device, requires_grad=True)
    (x + y).sum().backward()
    np.testing.assert_allclose(_np(x.grad), np.ones((2, 3)))
    np.testing.assert_allclose(_np(y.grad), np.full(3, 2.0))


@pytest.mark.parametrize("device", DEVICES)
def test_backward_matmul(device):
    """BASELINE config 1's path at the reference test's size (tests/test_backward.py:41-55)."""
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((2, 3)), rng.standard_normal((3, 4))
    x, w = pdn.Tensor(a, device=device, requires_grad=True), pdn.Tensor(b, device=device, requires_grad=True)
    pdn.matmul(x, w).sum().backward()
    np.testing.assert_allclose(_np(x.grad), np.ones((2, 4)) @ b.T, rtol=1e-6)
    np.testing.assert_allclose(_np(w.grad), a.T @ np.ones((2, 4)), rtol=1e-6)


@pytest.mark.parametrize("device", DEVICES)
def test_backward_retain_graph_and_errors(device):
    x = pdn.Tensor(2.0, device=device, requires_grad=True)
    y = x * x
    y.backward(retain_graph=True)
    np.testing.assert_allclose(_np(x.grad), 4.0)
    y.backward()
    np.testing.assert_allclose(_np(x.grad), 8.0)
    with pytest.raises(ValueError):
        y.backward()  # graph was freed
    v = pdn.Tensor(np.array([1.0, 2.0]), device=device, requires_grad=True)
    with pytest.raises(ValueError, match="scalar"):
        v.backward()
    with pytest.raises(ValueError):
        pdn.Tensor(1.0, device=device).backward()
    with pytest.raises(TypeError):
        pdn.Tensor(np.arange(3), device=device, requires_grad=True)
    with pytest.raises(ValueError):
        v += 1.0  # in-place on a grad-requiring tensor


@pytest.mark.parametrize("device", DEVICES)
def test_backward_concat_mean_slice_max(device):
    rng = np.random.default_rng(1)
    a, b = rng.standard_normal((2, 3)), rng.standard_normal((2, 2))
    x, y = pdn.Tensor(a, device=device, requires_grad=True), pdn.Tensor(b, device=device, requires_grad=True)
    w = rng.standard_normal((2, 5))
    (pdn.concat([x, y], axis=1) * pdn.Tensor(w, device=device)).sum().backward()
    np.testing.assert_allclose(_np(x.grad), w[:, :3])
    np.testing.assert_allclose(_np(y.grad), w[:, 3:])
    z = pdn.Tensor(a, device=device, requires_grad=True)
    pdn.mean(z, axis=1, keepdims=True).sum().backward()
    np.testing.assert_allclose(_np(z.grad), np.full((2, 3), 1 / 3))
    # duplicate fancy indices: last write wins, not accumulated (reference tensor.py:937-940)
    e = pdn.Tensor(rng.standard_normal((5, 2)), device=device, requires_grad=True)
    g = np.array([[1., 2.], [3., 4.], [5., 6.]])
    (e[np.array([1, 1, 3])] * pdn.Tensor(g, device=device)).sum().backward()
    exp = np.zeros((5, 2))
    exp[1], exp[3] = g[1], g[2]
    np.testing.assert_allclose(_np(e.grad), exp)
    # every tied maximum receives the full gradient (reference tensor.py:741-747)
    m = pdn.Tensor(np.array([[1., 3., 3.], [2., 0., 2.]]), device=device, requires_grad=True)
    m.max(axis=1).sum().backward()
    np.testing.assert_allclose(_np(m.grad), [[0, 1, 1], [1, 0, 1]])
    r = pdn.Tensor(np.array([-1., 0., 2.]), device=device, requires_grad=True)
    pdn.maximum(0., r).sum().backward()
    np.testing.assert_allclose(_np(r.grad), [0, 1, 1])  # relu'(0) = 1


@pytest.mark.parametrize("device", DEVICES)
def test_grad_dtype_quirk_and_nograd(device):
    x = pdn.Tensor(np.ones((2, 2), np.float32), device=device, requires_grad=True)
    assert x.grad.dtype == np.float64  # reference tensor.py:90
    p = pdn.Tensor(np.ones((2, 2), np.float32), dtype=np.float32, device=device, requires_grad=True)
    assert p.grad.dtype == np.float32
    with pdn.no_grad():
        y = p * 2
    assert not y.requires_grad
    assert (p * 2).requires_grad
    np.testing.assert_array_equal(_np(p.grad), 0)


@pytest.mark.parametrize("device", DEVICES)
def test_matmul_backward_shapes(device):
    """1-D promotion and broadcast-batch un-broadcasting of matmul grads (reference tensor.py:661-676, 360-370)."""
    rng = np.random.default_rng(3)
    cases = [((4, ), (4, 3)), ((2, 4), (4, )), ((4, ), (4, )), ((3, 2, 4), (4, 5)), ((3, 1, 2, 4), (1, 6, 4, 5)), ((2, 4), (3, 4, 5))]
    for sa, sb in cases:
        a, b = rng.standard_normal(sa), rng.standard_normal(sb)
        x, y = pdn.Tensor(a, device=device, requires_grad=True), pdn.Tensor(b, device=device, requires_grad=True)
        out = pdn.matmul(x, y)
        w = rng.standard_normal(out.shape)
        (out * pdn.Tensor(w, device=device)).sum().backward()
        eps = 1e-6
        for arr, t in ((a, x), (b, y)):
            num = np.zeros_like(arr)
            it = np.nditer(arr, flags=["multi_index"])
            for _ in it:
                i = it.multi_index
                old = arr[i]
                arr[i] = old + eps
                fp = (np.matmul(a, b) * w).sum()
                arr[i] = old - eps
                fm = (np.matmul(a, b) * w).sum()
                arr[i] = old
                num[i] = (fp - fm) / (2 * eps)
            np.testing.assert_allclose(_np(t.grad), num, rtol=1e-5, atol=1e-6)
