"""Tensor factories (reference pydynet/special.py:16-96). Random draws always come from the host ``np.random``
stream and are then moved to the device, so seeded models initialise bit-identically to the reference."""
import numpy as np

from .core import Tensor


def zeros(shape, dtype=None, device=None, requires_grad=False):
    return Tensor(np.zeros(shape), dtype=dtype, device=device, requires_grad=requires_grad)


def ones(shape, dtype=None, device=None, requires_grad=False):
    return Tensor(np.ones(shape), dtype=dtype, device=device, requires_grad=requires_grad)


def randn(*shape, dtype=None, device=None, requires_grad=False):
    return Tensor(np.random.randn(*shape), dtype=dtype, device=device, requires_grad=requires_grad)


def rand(*shape, dtype=None, device=None, requires_grad=False):
    return Tensor(np.random.rand(*shape), dtype=dtype, device=device, requires_grad=requires_grad)


def uniform(low: float, high: float, shape=None, dtype=None, device=None, requires_grad=False):
    return Tensor(np.random.uniform(low, high, size=shape), dtype=dtype, device=device, requires_grad=requires_grad)


def empty(shape, dtype=None, device=None, requires_grad=False):
    return Tensor(np.empty(shape, dtype=dtype), dtype=dtype, device=device, requires_grad=requires_grad)
