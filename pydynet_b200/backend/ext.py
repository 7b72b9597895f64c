"""Fused device helpers used by operator bodies on cuda tensors (twin of core/_host_ext.py)."""
from . import lib as L
from .array import ndarray, ternary, _unary, gemm_into, _reshape_strides, _prod


def eq_mul(a, b, g):
    """(a == b) * g in one pass — max/min/maximum/minimum grads (reference tensor.py:741-747, 812-823)."""
    return ternary(L.T_EQ_MUL, a, b, g)


def sigmoid(x):
    return _unary(L.SIGMOID, x)


def tanh(x):
    return _unary(L.TANH, x)


def sigmoid_grad(out, g):
    return ternary(L.T_SIGMOID_GRAD, out, g, g)


def tanh_grad(out, g):
    return ternary(L.T_TANH_GRAD, out, g, g)


def _flat2d(a: ndarray) -> ndarray:
    rows = _prod(a.shape[:-1])
    st = _reshape_strides(a.shape, a.estrides, (rows, a.shape[-1]))
    if st is None:
        a = a.copy()
        st = (a.shape[-1], 1)
    return a._view((rows, a.shape[-1]), st)


def matmul_dB(a, g, b_shape):
    """Aᵀ @ g (reference tensor.py:672-676). For (…, M, K) activations times a (K, N) weight the reference builds a
    per-batch dW and lets the engine sum it; here the GEMM contracts over all leading dims directly."""
    if len(b_shape) == 2 and a.ndim > 2 and g.ndim == a.ndim and a.shape[:-1] == g.shape[:-1]:
        a2, g2 = _flat2d(a), _flat2d(g)
        return gemm_into(None, a2.swapaxes(0, 1), g2)
    return a.swapaxes(-1, -2) @ g
