"""NumPy forms of the fused helpers used by operator bodies when a Tensor lives on the *cpu* device
(``Device.xp is numpy``, like the reference).  The cuda device never reaches this module: its twin is
pydynet_b200/backend/ext.py, which launches libpdn_b200.so kernels."""
import numpy as np


def eq_mul(a, b, g):
    return (a == b) * g


def sigmoid(x):
    # piecewise, overflow-safe (reference tensor.py:996-1002)
    out = np.empty_like(x)
    pos = x > 0
    out[pos] = 1 / (1 + np.exp(-x[pos]))
    out[~pos] = 1 - 1 / (1 + np.exp(x[~pos]))
    return out


def tanh(x):
    out = np.empty_like(x)
    pos = x > 0
    out[pos] = 2 / (1 + np.exp(-2 * x[pos])) - 1
    out[~pos] = 1 - 2 / (1 + np.exp(2 * x[~pos]))
    return out


def sigmoid_grad(out, g):
    return out * (1 - out) * g


def tanh_grad(out, g):
    return (1 - out**2) * g


def matmul_dB(a, g, b_shape):
    """Aᵀ @ g; when B is a plain matrix the contraction runs over all leading dims at once."""
    if len(b_shape) == 2 and a.ndim > 2 and g.ndim == a.ndim:
        return a.reshape(-1, a.shape[-1]).T @ g.reshape(-1, g.shape[-1])
    return a.swapaxes(-1, -2) @ g
