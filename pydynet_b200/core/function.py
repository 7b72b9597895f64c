"""Composite tensor helpers built from the operators of core/tensor.py (surface of reference
pydynet/core/function.py:4-259: sqrt, square, v/h/d/split, unsqueeze, squeeze)."""
from __future__ import annotations

import numbers

from .tensor import Tensor, reshape, swapaxes


def sqrt(x) -> Tensor:
    return x**0.5


def square(x) -> Tensor:
    return x * x


def _cut_points(total: int, indices_or_sections):
    """np.split's division rule: an int must divide the axis evenly; a sequence lists the cut positions."""
    if isinstance(indices_or_sections, numbers.Integral):
        n = int(indices_or_sections)
        if n <= 0:
            raise ValueError('number sections must be larger than 0.')
        assert total % n == 0, 'array split does not result in an equal division'
        step = total // n
        return [i * step for i in range(n + 1)]
    return [0] + [int(i) for i in indices_or_sections] + [total]


def split(x, indices_or_sections, axis: int = 0) -> list[Tensor]:
    if not isinstance(x, Tensor):
        x = Tensor(x)
    axis = axis % x.ndim
    pts = _cut_points(x.shape[axis], indices_or_sections)
    lead = (slice(None), ) * axis
    return [x[lead + (slice(a, b), )] for a, b in zip(pts[:-1], pts[1:])]


def vsplit(x, indices_or_sections) -> list[Tensor]:
    return split(x, indices_or_sections, 0)


def hsplit(x, indices_or_sections) -> list[Tensor]:
    return split(x, indices_or_sections, 1)


def dsplit(x, indices_or_sections) -> list[Tensor]:
    return split(x, indices_or_sections, 2)


def _axes(axis, ndim: int) -> tuple[int, ...]:
    if isinstance(axis, numbers.Integral):
        axis = (axis, )
    out = []
    for a in axis:
        a = int(a)
        if not -ndim <= a < ndim:
            raise ValueError(f"axis {a} is out of bounds for array of dimension {ndim}")
        a %= ndim
        if a in out:
            raise ValueError("repeated axis")
        out.append(a)
    return tuple(out)


def unsqueeze(x: Tensor, axis) -> Tensor:
    """np.expand_dims as a reshape node."""
    if isinstance(axis, numbers.Integral):
        axis = (axis, )
    nd = x.ndim + len(axis)
    where = _axes(axis, nd)
    it = iter(x.shape)
    return reshape(x, tuple(1 if i in where else next(it) for i in range(nd)))


def squeeze(x: Tensor, axis=None) -> Tensor:
    if axis is None:
        drop = tuple(i for i, s in enumerate(x.shape) if s == 1)
    else:
        drop = _axes(axis, x.ndim)
        for a in drop:
            if x.shape[a] != 1:
                raise ValueError("cannot select an axis to squeeze out which has size not equal to one")
    return reshape(x, tuple(s for i, s in enumerate(x.shape) if i not in drop))
