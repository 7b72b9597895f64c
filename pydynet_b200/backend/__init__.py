"""``xp`` namespace of the B200 backend — the module object ``Device.xp`` returns for cuda devices.

It stands where ``cupy`` stands in the reference (reference pydynet/cuda.py:90-91) and offers the
array-module surface the reference actually uses: constructors, elementwise math,
maximum/minimum, concatenate/expand_dims/atleast_2d/broadcast_to, ``random`` (host NumPy draws moved to
the device, so RNG order matches the reference's CPU-side initialisers, reference nn/init.py:29-38).
"""
from __future__ import annotations

import numbers

import numpy as np

from . import lib as L
from .array import (ndarray, matmul, gemm_into, ternary, _binary, _unary, _reduce, _to_dev, _code, _arr, _prod,
                    _contig_strides, _scatter, _apply_basic)

bool_ = np.bool_
float16, float32, float64, int32, int64, intp = np.float16, np.float32, np.float64, np.int32, np.int64, np.intp
issubdtype = np.issubdtype
floating, integer = np.floating, np.integer
newaxis = None
inf = np.inf


def array(obj, dtype=None, copy=True) -> ndarray:
    if isinstance(obj, ndarray):
        if dtype is not None and np.dtype(dtype) != obj.dtype:
            return obj.astype(dtype)
        return obj.copy() if copy else obj
    return ndarray.from_host(obj, dtype=dtype)


def asarray(obj, dtype=None) -> ndarray:
    return array(obj, dtype=dtype, copy=None)


def asnumpy(a) -> np.ndarray:
    return a.get() if isinstance(a, ndarray) else np.asarray(a)


def empty(shape, dtype=np.float64) -> ndarray:
    return ndarray.empty(shape, dtype)


def full(shape, value, dtype=None) -> ndarray:
    if dtype is None:
        dtype = np.asarray(value).dtype
    out = ndarray.empty(shape, dtype)
    out.fill(value)
    return out


def zeros(shape, dtype=np.float64) -> ndarray:
    return full(shape, 0, np.float64 if dtype is None else dtype)


def ones(shape, dtype=np.float64) -> ndarray:
    return full(shape, 1, np.float64 if dtype is None else dtype)


def zeros_like(a, dtype=None) -> ndarray:
    return zeros(a.shape, dtype or a.dtype)


def ones_like(a, dtype=None) -> ndarray:
    return ones(a.shape, dtype or a.dtype)


def empty_like(a, dtype=None) -> ndarray:
    return ndarray.empty(a.shape, dtype or a.dtype)


def exp(a): return _unary(L.EXP, a)
def log(a): return _unary(L.LOG, a)
def abs(a): return _unary(L.ABS, a)
def sign(a): return _unary(L.SIGN, a)
def sqrt(a): return _unary(L.SQRT, a)
def square(a): return _unary(L.SQUARE, a)
def negative(a): return _unary(L.NEG, a)
def maximum(a, b): return _binary(L.MAXIMUM, a, b)
def minimum(a, b): return _binary(L.MINIMUM, a, b)
def add(a, b): return _binary(L.ADD, a, b)
def subtract(a, b): return _binary(L.SUB, a, b)
def multiply(a, b): return _binary(L.MUL, a, b)
def divide(a, b): return _binary(L.DIV, a, b, true_div=True)
def power(a, b): return _binary(L.POW, a, b)
def sum(a, axis=None, keepdims=False): return _reduce(L.R_SUM, a, axis, keepdims)
def mean(a, axis=None, keepdims=False): return _reduce(L.R_MEAN, a, axis, keepdims)
def max(a, axis=None, keepdims=False): return _reduce(L.R_MAX, a, axis, keepdims)
def min(a, axis=None, keepdims=False): return _reduce(L.R_MIN, a, axis, keepdims)
def argmax(a, axis=None, keepdims=False): return _reduce(L.R_ARGMAX, a, axis, keepdims)
def argmin(a, axis=None, keepdims=False): return _reduce(L.R_ARGMIN, a, axis, keepdims)


def where(cond, a, b) -> ndarray:
    a, b = _to_dev(a), _to_dev(b)
    dt = np.result_type(a.dtype, b.dtype)
    return ternary(L.T_WHERE, _to_dev(cond).astype(dt), a, b)


def expand_dims(a: ndarray, axis) -> ndarray:
    if isinstance(axis, numbers.Integral):
        axis = (int(axis), )
    nd = a.ndim + len(axis)
    axis = sorted(ax % nd for ax in axis)
    shape, strides = list(a.shape), list(a.estrides)
    for ax in axis:
        shape.insert(ax, 1)
        strides.insert(ax, 0)
    return a._view(shape, strides)


def squeeze(a: ndarray, axis=None) -> ndarray:
    if axis is None:
        axis = tuple(i for i, s in enumerate(a.shape) if s == 1)
    elif isinstance(axis, numbers.Integral):
        axis = (int(axis) % a.ndim, )
    else:
        axis = tuple(int(x) % a.ndim for x in axis)
    for ax in axis:
        if a.shape[ax] != 1:
            raise ValueError("cannot select an axis to squeeze out which has size not equal to one")
    keep = [i for i in range(a.ndim) if i not in axis]
    return a._view([a.shape[i] for i in keep], [a.estrides[i] for i in keep])


def atleast_2d(a: ndarray) -> ndarray:
    if a.ndim >= 2:
        return a
    if a.ndim == 1:
        return a._view((1, ) + a.shape, (0, ) + a.estrides)
    return a._view((1, 1), (0, 0))


def broadcast_to(a: ndarray, shape) -> ndarray:
    return a.broadcast_to(shape)


def reshape(a: ndarray, shape) -> ndarray:
    return a.reshape(shape)


def transpose(a: ndarray, axes=None) -> ndarray:
    return a.transpose(axes)


def swapaxes(a: ndarray, a1, a2) -> ndarray:
    return a.swapaxes(a1, a2)


def ascontiguousarray(a: ndarray) -> ndarray:
    return a.ascontiguous()


def concatenate(arrays, axis=0) -> ndarray:
    arrays = [_to_dev(a) for a in arrays]
    if not arrays:
        raise ValueError("need at least one array to concatenate")
    nd = arrays[0].ndim
    if nd == 0:
        raise ValueError("zero-dimensional arrays cannot be concatenated")
    axis %= nd
    dt = np.result_type(*[a.dtype for a in arrays])
    shape = list(arrays[0].shape)
    for a in arrays[1:]:
        if a.ndim != nd or any(a.shape[i] != shape[i] for i in range(nd) if i != axis):
            raise ValueError("all the input array dimensions except for the concatenation axis must match exactly")
    shape[axis] = builtins_sum(a.shape[axis] for a in arrays)
    out = ndarray.empty(shape, dt)
    pos = 0
    for a in arrays:
        n = a.shape[axis]
        if n:
            sub_shape = list(shape)
            sub_shape[axis] = n
            dst = out._view(sub_shape, out.estrides, pos * out.estrides[axis])
            a._copy_into(dst)
        pos += n
    return out


def stack(arrays, axis=0) -> ndarray:
    return concatenate([expand_dims(_to_dev(a), axis) for a in arrays], axis)


def pad(a: ndarray, pad_width, mode="constant") -> ndarray:
    assert mode == "constant"
    shape = [s + lo + hi for s, (lo, hi) in zip(a.shape, pad_width)]
    out = zeros(shape, a.dtype)
    inner = out._view(a.shape, out.estrides, builtins_sum(lo * st for (lo, _), st in zip(pad_width, out.estrides)))
    a._copy_into(inner)
    return out


def add_at(dst: ndarray, key, values):
    """np.add.at(dst, key, values) for integer-array keys."""
    view, adv = _apply_basic(dst, key)
    if adv is None:
        view += values
    else:
        _scatter(view, adv, values, accumulate=True)


import builtins as _b  # noqa: E402

builtins_sum = _b.sum


class _Random:
    """``xp.random``: draws come from host ``np.random`` (same global stream, same order as the reference's
    initialisers, reference nn/init.py:31,37 and special.py:48-96) and are then moved to the device."""

    @staticmethod
    def uniform(low=0.0, high=1.0, size=None):
        return ndarray.from_host(np.random.uniform(low, high, size))

    @staticmethod
    def normal(loc=0.0, scale=1.0, size=None):
        return ndarray.from_host(np.random.normal(loc, scale, size))

    @staticmethod
    def rand(*shape):
        return ndarray.from_host(np.random.rand(*shape))

    @staticmethod
    def randn(*shape):
        return ndarray.from_host(np.random.randn(*shape))

    @staticmethod
    def seed(s):
        np.random.seed(s)


random = _Random()

from . import ext  # noqa: E402,F401
