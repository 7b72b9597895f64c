"""pydynet_b200 — PyDyNet's Tensor / autograd / nn surface on a hand-written sm_100a backend.

Same flat exports as reference pydynet/__init__.py:1-17.
"""
from .core import (Tensor, add, sub, mul, div, pow, matmul, abs, sum, mean, min, max, argmax, argmin, maximum, minimum, exp,
                   log, sign, reshape, transpose, swapaxes, concat, sigmoid, tanh, sqrt, square, vsplit, hsplit, dsplit,
                   split, unsqueeze, squeeze)
from .special import zeros, ones, rand, randn, empty, uniform
from .cuda import Device
from .autograd import enable_grad, no_grad
from . import cuda, autograd, core, special
from . import nn, optim

__all__ = [
    "Tensor", "add", "sub", "mul", "div", "pow", "matmul", "abs", "sum", "mean", "min", "max", "argmax", "argmin", "maximum",
    "minimum", "exp", "log", "sign", "reshape", "transpose", "swapaxes", "concat", 'sigmoid', 'tanh', "sqrt", "square",
    "vsplit", "hsplit", "dsplit", "split", "unsqueeze", "squeeze", "zeros", "ones", "rand", "randn", "empty", "uniform",
    "Device", "enable_grad", "no_grad"
]
