"""Device ndarray of the B200 backend: the object that sits where a ``cupy.ndarray`` sits in the reference.

The reference does all Tensor math as ``self.xp.<fn>(...)`` / ndarray operators on ``tensor.data``
(reference pydynet/core/tensor.py:80, 512-529; surface listed in SURVEY.md §8b).  This class provides that
surface — NumPy broadcasting, dtype promotion, zero-copy views for reshape/transpose/basic slices (the
reference relies on view aliasing, e.g. the in-place KV-cache writes of llm/llama/model.py:106-107 and
``Parameter`` sharing its source buffer, nn/parameter.py:7-13) — on top of the C ABI in
include/pdn_b200.h.  Host NumPy is used here only for shape/stride/dtype *metadata* arithmetic and for
H2D/D2H staging; all element math runs in libpdn_b200.so kernels.
"""
from __future__ import annotations

import ctypes as C
import os
import math
import numbers

import numpy as np

from . import lib as L

_DT = {
    np.dtype(np.float32): L.F32, np.dtype(np.float64): L.F64, np.dtype(np.float16): L.F16,
    np.dtype(np.int64): L.I64, np.dtype(np.int32): L.I32, np.dtype(np.bool_): L.BOOL, np.dtype(np.uint8): L.U8,
}
_FLOATS = (np.dtype(np.float16), np.dtype(np.float32), np.dtype(np.float64))
_I64A = C.c_int64 * 8
_F32 = np.dtype(np.float32)
_ONE3, _ZERO3 = (C.c_int64 * 3)(1, 1, 1), (C.c_int64 * 3)(0, 0, 0)
# bumped by every USER-LEVEL in-place write to a device buffer (setitem, fill, += ...): inference plans (nn/_plans.py) compare it
# against the value they saw last and re-check their weight signatures only when it moved
WRITE_EPOCH = [0]
IMPLICIT_NUMPY = os.environ.get("PDN_IMPLICIT_NUMPY") == "1"


def _code(dt) -> int:
    try:
        return _DT[np.dtype(dt)]
    except KeyError:
        raise TypeError(f"dtype {dt} is not supported by the B200 backend") from None


def _arr(vals):
    a = _I64A()
    for i, v in enumerate(vals):
        a[i] = v
    return a


class _Buffer:
    """Owns one allocation of the caching allocator; freed when the last view dies."""
    __slots__ = ("ptr", "nbytes", "version")

    def __init__(self, nbytes: int):
        p = C.c_void_p()
        L.call("pdn_malloc", C.byref(p), max(int(nbytes), 1))
        self.ptr = p.value or 0
        self.nbytes = nbytes
        self.version = 0  # bumped by every in-place write; derived caches (pre-packed GEMM weights) key on it

    def __del__(self):
        try:
            if self.ptr and L._lib is not None:
                L._lib.pdn_free(self.ptr)
        except Exception:
            pass


def _contig_strides(shape):
    st, acc = [], 1
    for s in reversed(shape):
        st.append(acc)
        acc *= max(s, 1) if s != 0 else 1
    return tuple(reversed(st))


def _prod(shape):
    n = 1
    for s in shape:
        n *= s
    return n


class ndarray:
    __slots__ = ("buf", "ptr", "shape", "estrides", "dtype", "size", "__weakref__")
    __array_priority__ = 1000  # numpy scalars/arrays defer to us in mixed operators

    def __init__(self, buf, ptr, shape, estrides, dtype):
        self.buf = buf
        self.ptr = ptr
        self.shape = tuple(map(int, shape))
        self.estrides = tuple(map(int, estrides))
        self.dtype = dtype if isinstance(dtype, np.dtype) else np.dtype(dtype)
        n = 1
        for d in self.shape:
            n *= d
        self.size = n

    # ------------------------------------------------------------------ construction ------------
    @staticmethod
    def empty(shape, dtype=np.float32) -> "ndarray":
        if isinstance(shape, numbers.Integral):
            shape = (int(shape), )
        shape = tuple(map(int, shape))
        if not isinstance(dtype, np.dtype):
            dtype = np.dtype(dtype)
        if dtype not in _DT:
            _code(dtype)  # raises the TypeError
        # (one pass: contiguous strides and the element count together; the object is filled without re-validating what was just built)
        st, n = [], 1
        for d in reversed(shape):
            st.append(n)
            n *= d if d != 0 else 1
        size = 1
        for d in shape:
            size *= d
        buf = _Buffer(size * dtype.itemsize)
        out = ndarray.__new__(ndarray)
        out.buf, out.ptr, out.shape, out.estrides, out.dtype, out.size = buf, buf.ptr, shape, tuple(reversed(st)), dtype, size
        return out

    @staticmethod
    def from_host(a, dtype=None) -> "ndarray":
        a = np.asarray(a, dtype=dtype)
        if a.dtype not in _DT:
            if np.issubdtype(a.dtype, np.integer):
                a = a.astype(np.int64)
            elif np.issubdtype(a.dtype, np.floating):
                a = a.astype(np.float64)
            else:
                raise TypeError(f"cannot move dtype {a.dtype} to the device")
        a = np.require(a, requirements='C')  # (np.ascontiguousarray would promote 0-d to 1-d)
        out = ndarray.empty(a.shape, a.dtype)
        if a.size:
            L.call("pdn_memcpy_h2d", out.ptr, a.ctypes.data, a.nbytes)
        return out

    # ------------------------------------------------------------------ metadata ----------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def itemsize(self):
        return self.dtype.itemsize

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def strides(self):
        return tuple(s * self.dtype.itemsize for s in self.estrides)

    @property
    def T(self):
        return self.transpose()

    @property
    def is_contiguous(self) -> bool:
        exp = 1
        for s, st in zip(reversed(self.shape), reversed(self.estrides)):
            if s == 1:
                continue
            if st != exp:
                return False
            exp *= s
        return True

    def __len__(self):
        if not self.shape:
            raise TypeError("len() of unsized object")
        return self.shape[0]

    def _view(self, shape, estrides, offset_elems=0) -> "ndarray":
        return ndarray(self.buf, self.ptr + offset_elems * self.dtype.itemsize, shape, estrides, self.dtype)

    # ------------------------------------------------------------------ host transfer -----------
    def get(self) -> np.ndarray:
        """D2H copy (synchronises) — cupy's ``.get()`` used by reference Tensor.numpy (tensor.py:385-390)."""
        src = self if self.is_contiguous else self.copy()
        out = np.empty(self.shape, dtype=self.dtype)
        L.call("pdn_memcpy_d2h", out.ctypes.data, src.ptr, out.nbytes)
        return out

    def item(self):
        if self.size != 1:
            raise ValueError("can only convert an array of size 1 to a Python scalar")
        return self.get().item()

    def tolist(self):
        return self.get().tolist()

    def __float__(self):
        return float(self.item())

    def __int__(self):
        return int(self.item())

    def __bool__(self):
        if self.size != 1:
            raise ValueError("The truth value of an array with more than one element is ambiguous.")
        return bool(self.item())

    def __repr__(self):
        return "device" + repr(self.get())

    def __str__(self):
        return str(self.get())

    def __array__(self, dtype=None, copy=None):
        # like CuPy, a device array does not silently turn into a host array (a hidden synchronising copy); PDN_IMPLICIT_NUMPY=1 opts
        # in — tests/test_variant_b_gpu.py uses it: the reference's own test-suite hands raw ``.grad`` arrays to np.testing
        if IMPLICIT_NUMPY:
            out = self.get()
            return out if dtype is None else out.astype(dtype)
        raise TypeError("implicit conversion of a device array to NumPy is not allowed; call .get()")

    # ------------------------------------------------------------------ copies / casts ----------
    def _copy_into(self, dst: "ndarray"):
        """dst[...] = self (same shape), strided, with cast."""
        assert dst.shape == self.shape, (dst.shape, self.shape)
        dst.buf.version += 1
        L.call("pdn_copy", self.ptr, _code(self.dtype), dst.ptr, _code(dst.dtype), len(self.shape), _arr(self.shape),
               _arr(self.estrides), _arr(dst.estrides))

    def copy(self) -> "ndarray":
        out = ndarray.empty(self.shape, self.dtype)
        self._copy_into(out)
        return out

    def astype(self, dtype, copy=True) -> "ndarray":
        dtype = np.dtype(dtype)
        if dtype == self.dtype and not copy:
            return self
        out = ndarray.empty(self.shape, dtype)
        self._copy_into(out)
        return out

    def ascontiguous(self) -> "ndarray":
        return self if self.is_contiguous else self.copy()

    def fill(self, value):
        self.buf.version += 1
        WRITE_EPOCH[0] += 1
        L.call("pdn_fill", self.ptr, _code(self.dtype), len(self.shape), _arr(self.shape), _arr(self.estrides), float(value))

    # ------------------------------------------------------------------ views -------------------
    def reshape(self, *shape) -> "ndarray":
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        shape = [int(s) for s in shape]
        if shape.count(-1) > 1:
            raise ValueError("can only specify one unknown dimension")
        if -1 in shape:
            known = -_prod(shape)
            if known == 0 or self.size % known:
                raise ValueError(f"cannot reshape array of size {self.size} into shape {tuple(shape)}")
            shape[shape.index(-1)] = self.size // known
        if _prod(shape) != self.size:
            raise ValueError(f"cannot reshape array of size {self.size} into shape {tuple(shape)}")
        st = _reshape_strides(self.shape, self.estrides, shape)
        if st is None:  # NumPy would copy here too
            c = self.copy()
            return c._view(shape, _contig_strides(shape))
        return self._view(shape, st)

    def transpose(self, *axes) -> "ndarray":
        if len(axes) == 1 and (axes[0] is None or isinstance(axes[0], (tuple, list))):
            axes = axes[0]
        if axes is None or len(axes) == 0:
            axes = tuple(reversed(range(self.ndim)))
        axes = tuple(a % self.ndim if self.ndim else 0 for a in axes)
        if sorted(axes) != list(range(self.ndim)):
            raise ValueError("axes don't match array")
        return self._view([self.shape[a] for a in axes], [self.estrides[a] for a in axes])

    def swapaxes(self, a1, a2) -> "ndarray":
        ax = list(range(self.ndim))
        a1 %= self.ndim
        a2 %= self.ndim
        ax[a1], ax[a2] = ax[a2], ax[a1]
        return self.transpose(ax)

    def broadcast_to(self, shape) -> "ndarray":
        shape = tuple(map(int, shape))
        if shape == self.shape:
            return self
        nd = len(shape)
        if nd < self.ndim:
            raise ValueError("cannot broadcast to fewer dimensions")
        sh = (1, ) * (nd - self.ndim) + self.shape
        st = (0, ) * (nd - self.ndim) + self.estrides
        out_st = []
        for s_from, s_to, stv in zip(sh, shape, st):
            if s_from == s_to:
                out_st.append(stv)
            elif s_from == 1:
                out_st.append(0)
            else:
                raise ValueError(f"operands could not be broadcast together with shapes {self.shape} {shape}")
        return self._view(shape, out_st)

    # ------------------------------------------------------------------ indexing ----------------
    def __getitem__(self, key) -> "ndarray":
        view, adv = _apply_basic(self, key)
        if adv is None:
            return view
        return _gather(view, adv)

    def __setitem__(self, key, value):
        WRITE_EPOCH[0] += 1
        view, adv = _apply_basic(self, key)
        if adv is None:
            _assign(view, value)
        else:
            _scatter(view, adv, value, accumulate=False)

    # ------------------------------------------------------------------ arithmetic --------------
    def __add__(self, o): return _binary(L.ADD, self, o)
    def __radd__(self, o): return _binary(L.ADD, o, self)
    def __sub__(self, o): return _binary(L.SUB, self, o)
    def __rsub__(self, o): return _binary(L.SUB, o, self)
    def __mul__(self, o): return _binary(L.MUL, self, o)
    def __rmul__(self, o): return _binary(L.MUL, o, self)
    def __truediv__(self, o): return _binary(L.DIV, self, o, true_div=True)
    def __rtruediv__(self, o): return _binary(L.DIV, o, self, true_div=True)
    def __pow__(self, o): return _binary(L.POW, self, o)
    def __rpow__(self, o): return _binary(L.POW, o, self)
    def __neg__(self): return _unary(L.NEG, self)
    def __pos__(self): return self
    def __abs__(self): return _unary(L.ABS, self)
    def __matmul__(self, o): return matmul(self, o)
    def __rmatmul__(self, o): return matmul(o, self)
    def __eq__(self, o): return _binary(L.EQ, self, o)
    def __ne__(self, o): return _binary(L.NE, self, o)
    def __lt__(self, o): return _binary(L.LT, self, o)
    def __le__(self, o): return _binary(L.LE, self, o)
    def __gt__(self, o): return _binary(L.GT, self, o)
    def __ge__(self, o): return _binary(L.GE, self, o)
    __hash__ = None

    def __iadd__(self, o): return _binary(L.ADD, self, o, out=self)
    def __isub__(self, o): return _binary(L.SUB, self, o, out=self)
    def __imul__(self, o): return _binary(L.MUL, self, o, out=self)
    def __itruediv__(self, o): return _binary(L.DIV, self, o, out=self, true_div=True)

    def __imatmul__(self, o):
        r = matmul(self, o)
        if r.shape != self.shape:
            raise ValueError("in-place matmul changes the shape")
        r._copy_into(self)
        WRITE_EPOCH[0] += 1
        return self

    # ------------------------------------------------------------------ reductions --------------
    def sum(self, axis=None, keepdims=False, dtype=None): return _reduce(L.R_SUM, self, axis, keepdims)
    def mean(self, axis=None, keepdims=False, dtype=None): return _reduce(L.R_MEAN, self, axis, keepdims)
    def max(self, axis=None, keepdims=False): return _reduce(L.R_MAX, self, axis, keepdims)
    def min(self, axis=None, keepdims=False): return _reduce(L.R_MIN, self, axis, keepdims)
    def cumsum(self, axis=None):
        """Index arithmetic only (reference core/function.py:39,75,111: section boundaries of split): computed on the host."""
        return ndarray.from_host(np.cumsum(self.get(), axis=axis))

    def __index__(self):
        if self.size != 1 or not np.issubdtype(self.dtype, np.integer):
            raise TypeError("only integer scalar device arrays can be converted to an index")
        return int(self.get().reshape(()))

    def argmax(self, axis=None, keepdims=False): return _reduce(L.R_ARGMAX, self, axis, keepdims)
    def argmin(self, axis=None, keepdims=False): return _reduce(L.R_ARGMIN, self, axis, keepdims)


# ---------------------------------------------------------------------- helpers ------------------
def _reshape_strides(shape, strides, newshape):
    """NumPy's no-copy reshape rule; returns new element strides or None if a copy is needed."""
    old = [(s, st) for s, st in zip(shape, strides) if s != 1]
    if _prod(newshape) == 0:
        return _contig_strides(newshape)
    new_st = [0] * len(newshape)
    oi, ni, on, nn = 0, 0, len(old), len(newshape)
    while oi < on and ni < nn:
        op, np_ = old[oi][0], newshape[ni]
        oj, nj = oi + 1, ni + 1
        while op != np_:
            if np_ < op:
                np_ *= newshape[nj]
                nj += 1
            else:
                op *= old[oj][0]
                oj += 1
        for k in range(oi, oj - 1):  # old dims oi..oj-1 must be mutually contiguous
            if old[k][1] != old[k + 1][0] * old[k + 1][1]:
                return None
        st = old[oj - 1][1]
        for k in range(nj - 1, ni - 1, -1):
            new_st[k] = st
            st *= newshape[k]
        oi, ni = oj, nj
    for k in range(nn):  # trailing / interleaved size-1 dims
        if newshape[k] == 1 and new_st[k] == 0:
            new_st[k] = 1
    return tuple(new_st)


def _is_index_array(k):
    return isinstance(k, (ndarray, np.ndarray, list, range))


def _apply_basic(a: ndarray, key):
    """Applies the basic (view) part of an index expression; returns (view, adv) where adv is None or
    (first_axis, [index arrays]) for adjacent integer-array indices."""
    if not isinstance(key, tuple):
        key = (key, )
    if len(key) == 1 and isinstance(key[0], (bool, np.bool_)):
        # NumPy: a[True] is a[None], a[False] is an empty selection (the reference hits this through `t == 1`, which is
        # Python identity because Tensor defines no __eq__ — examples/pydynet/transformer.py:242-243)
        view = a._view((1, ) + a.shape, (0, ) + a.estrides)
        return (view, None) if key[0] else (view._view((0, ) + a.shape, (0, ) + a.estrides), None)
    # whole-array boolean mask
    if len(key) == 1 and ((isinstance(key[0], ndarray) and key[0].dtype == np.bool_) or
                          (isinstance(key[0], np.ndarray) and key[0].dtype == np.bool_)):
        m = key[0].get() if isinstance(key[0], ndarray) else key[0]
        if m.shape != a.shape[:m.ndim]:
            raise IndexError("boolean index did not match indexed array")
        nz = np.nonzero(m)
        return a, (0, [np.asarray(i, dtype=np.int64) for i in nz])
    has_arrays = any(_is_index_array(k) for k in key)
    n_real = sum(1 for k in key if k is not None and k is not Ellipsis)
    if n_real > a.ndim:
        raise IndexError(f"too many indices for array: array is {a.ndim}-dimensional, but {n_real} were indexed")
    if sum(1 for k in key if k is Ellipsis) > 1:
        raise IndexError("an index can only have a single ellipsis ('...')")
    expanded = []
    for k in key:
        if k is Ellipsis:
            expanded.extend([slice(None)] * (a.ndim - n_real))
        else:
            expanded.append(k)
    if not any(k is Ellipsis for k in key):
        expanded.extend([slice(None)] * (a.ndim - n_real))
    shape, strides, offset = [], [], 0
    adv_axes, adv_arrays = [], []
    dim = 0
    for k in expanded:
        if k is None:
            shape.append(1)
            strides.append(0)
            continue
        n, st = a.shape[dim], a.estrides[dim]
        if isinstance(k, slice):
            start, stop, step = k.indices(n)
            ln = len(range(start, stop, step))
            shape.append(ln)
            strides.append(st * step)
            offset += start * st if ln > 0 else 0
        elif isinstance(k, numbers.Integral) and not isinstance(k, bool) and not has_arrays:
            i = int(k)
            if i < -n or i >= n:
                raise IndexError(f"index {i} is out of bounds for axis {dim} with size {n}")
            offset += (i % n) * st
        else:  # integer array (or an int mixed with arrays -> 0-d array)
            if isinstance(k, numbers.Integral):
                k = np.asarray(int(k), dtype=np.int64)
            elif isinstance(k, (list, range)):
                k = np.asarray(k)
                if k.dtype == np.bool_:
                    k = np.nonzero(k)[0]
                if k.size == 0:
                    k = k.astype(np.int64)
            if isinstance(k, np.ndarray) and not np.issubdtype(k.dtype, np.integer):
                raise IndexError("arrays used as indices must be of integer (or boolean) type")
            if isinstance(k, ndarray) and not np.issubdtype(k.dtype, np.integer):
                raise IndexError("arrays used as indices must be of integer (or boolean) type")
            adv_axes.append(len(shape))
            adv_arrays.append(k)
            shape.append(n)
            strides.append(st)
        dim += 1
    view = a._view(shape, strides, offset)
    if not adv_axes:
        return view, None
    if adv_axes != list(range(adv_axes[0], adv_axes[0] + len(adv_axes))):
        raise NotImplementedError("integer-array indices separated by slices are not supported by the B200 backend")
    if len(adv_axes) > 4:
        raise NotImplementedError("more than 4 integer-array indices")
    return view, (adv_axes[0], adv_arrays)


def _index_args(view: ndarray, adv):
    first, arrays = adv
    K = len(arrays)
    bshape = np.broadcast_shapes(*[tuple(x.shape) for x in arrays])
    J = _prod(bshape)
    dev_idx = []
    for x in arrays:
        if isinstance(x, ndarray):
            x = x if x.dtype == np.int64 else x.astype(np.int64)
            x = x.broadcast_to(bshape).ascontiguous() if x.shape != tuple(bshape) else x.ascontiguous()
        else:
            x = ndarray.from_host(np.array(np.broadcast_to(np.asarray(x, dtype=np.int64), bshape), order='C'))
        dev_idx.append(x)
    ptrs = (C.c_void_p * 4)(*[x.ptr for x in dev_idx])
    outer_shape, outer_st = view.shape[:first], view.estrides[:first]
    inner_shape, inner_st = view.shape[first + K:], view.estrides[first + K:]
    if len(outer_shape) > 4 or len(inner_shape) > 4:
        raise NotImplementedError("index: more than 4 outer/inner dims")
    dims = view.shape[first:first + K]
    dst = view.estrides[first:first + K]
    return dev_idx, ptrs, K, dims, dst, J, tuple(bshape), outer_shape, outer_st, inner_shape, inner_st


def _gather(view: ndarray, adv) -> ndarray:
    keep, ptrs, K, dims, dst, J, bshape, os_, ost, is_, ist = _index_args(view, adv)
    out = ndarray.empty(tuple(os_) + bshape + tuple(is_), view.dtype)
    if out.size:
        L.call("pdn_index_gather", view.ptr, _code(view.dtype), out.ptr, K, ptrs, _arr(dims), _arr(dst), J, len(os_), _arr(os_),
               _arr(ost), len(is_), _arr(is_), _arr(ist))
    return out


def _scatter(view: ndarray, adv, value, accumulate: bool):
    keep, ptrs, K, dims, dst, J, bshape, os_, ost, is_, ist = _index_args(view, adv)
    tshape = tuple(os_) + bshape + tuple(is_)
    if not isinstance(value, ndarray):
        value = ndarray.from_host(np.broadcast_to(np.asarray(value, dtype=view.dtype), tshape))
    if value.dtype != view.dtype:
        value = value.astype(view.dtype)
    if value.shape != tshape:
        value = value.broadcast_to(tshape)
    value = value.ascontiguous()
    view.buf.version += 1
    if value.size:
        L.call("pdn_index_scatter", view.ptr, _code(view.dtype), value.ptr, K, ptrs, _arr(dims), _arr(dst), J, len(os_), _arr(os_),
               _arr(ost), len(is_), _arr(is_), _arr(ist), 1 if accumulate else 0)


def _assign(view: ndarray, value):
    """view[...] = value with NumPy broadcasting and casting."""
    if isinstance(value, ndarray):
        src = value
        if src.shape != view.shape:
            # NumPy allows extra leading 1-dims on the value
            while src.ndim > view.ndim and src.shape[0] == 1:
                src = src._view(src.shape[1:], src.estrides[1:])
            src = src.broadcast_to(view.shape)
        src._copy_into(view)
    elif isinstance(value, (numbers.Number, np.generic)) or (isinstance(value, np.ndarray) and value.ndim == 0):
        view.fill(float(value))
    else:
        h = np.asarray(value)
        src = ndarray.from_host(h if h.dtype in _DT else h.astype(view.dtype))
        _assign(view, src)


def _result_dtype(a, b):
    da = a.dtype if isinstance(a, (ndarray, np.ndarray, np.generic)) else a
    db = b.dtype if isinstance(b, (ndarray, np.ndarray, np.generic)) else b
    return np.result_type(da, db)


def _to_dev(x, dtype=None) -> ndarray:
    if isinstance(x, ndarray):
        return x if dtype is None or x.dtype == dtype else x.astype(dtype)
    return ndarray.from_host(np.asarray(x, dtype=dtype))


def _is_scalar(x):
    return isinstance(x, (numbers.Number, np.generic)) or (isinstance(x, np.ndarray) and x.ndim == 0)


def _binary(op, a, b, out: ndarray | None = None, true_div: bool = False) -> ndarray:
    cmp = op >= L.EQ
    # ---- scalar operand: kernel takes it as an argument (no H2D of a 0-d array)
    if _is_scalar(a) or _is_scalar(b):
        arr, sc, rev = (b, a, 1) if _is_scalar(a) else (a, b, 0)
        if isinstance(sc, np.ndarray):
            sc = sc[()]
        if isinstance(sc, np.generic) and not isinstance(sc, (np.floating, np.integer, np.bool_)):
            raise TypeError("unsupported scalar type")
        arr = _to_dev(arr)
        dt = np.result_type(arr.dtype, sc)
        if true_div and not np.issubdtype(dt, np.floating):
            dt = np.dtype(np.float64)
        if dt == np.bool_ and not cmp:
            dt = np.dtype(np.int64) if op in (L.ADD, L.SUB, L.MUL) else dt
        if out is not None:
            out.buf.version += 1
            WRITE_EPOCH[0] += 1
            dt_c = out.dtype  # in-place ops compute in the destination dtype
            x = arr if arr.dtype == dt_c else arr.astype(dt_c)
            L.call("pdn_ew_binary_scalar", op, _code(dt_c), x.ptr, float(sc), rev, out.ptr, len(out.shape), _arr(out.shape),
                   _arr(x.broadcast_to(out.shape).estrides), _arr(out.estrides))
            return out
        x = arr if arr.dtype == dt else arr.astype(dt)
        res = ndarray.empty(x.shape, np.bool_ if cmp else dt)
        if res.size:
            L.call("pdn_ew_binary_scalar", op, _code(dt), x.ptr, float(sc), rev, res.ptr, len(x.shape), _arr(x.shape),
                   _arr(x.estrides), _arr(res.estrides))
        return res
    a, b = _to_dev(a), _to_dev(b)
    dt = np.result_type(a.dtype, b.dtype)
    if true_div and not np.issubdtype(dt, np.floating):
        dt = np.dtype(np.float64)
    if dt == np.bool_ and not cmp:
        if op not in (L.ADD, L.MAXIMUM, L.MUL, L.MINIMUM):  # logical or / and
            raise TypeError("numpy boolean subtract/divide/power is not supported")
        r = _binary(op, a.astype(np.int32), b.astype(np.int32))
        return r.astype(np.bool_)
    if out is not None:
        out.buf.version += 1
        WRITE_EPOCH[0] += 1
        dt = out.dtype  # in-place ops compute in the destination dtype (same-kind casting)
        shape = out.shape
        if np.broadcast_shapes(a.shape, b.shape) != shape:
            raise ValueError(f"non-broadcastable output operand with shape {shape}")
    else:
        shape = np.broadcast_shapes(a.shape, b.shape)
    if a.dtype != dt:
        a = a.astype(dt)
    if b.dtype != dt:
        b = b.astype(dt)
    av, bv = a.broadcast_to(shape), b.broadcast_to(shape)
    res = out if out is not None else ndarray.empty(shape, np.bool_ if cmp else dt)
    if res.size:
        L.call("pdn_ew_binary", op, _code(dt), av.ptr, bv.ptr, res.ptr, len(shape), _arr(shape), _arr(av.estrides), _arr(bv.estrides),
               _arr(res.estrides))
    return res


def _unary(op, a, out=None) -> ndarray:
    a = _to_dev(a)
    if op not in (L.NEG, L.ABS, L.SIGN, L.SQUARE, L.RELU) and a.dtype not in _FLOATS:
        a = a.astype(np.float64)
    if a.dtype == np.bool_:
        raise TypeError("unary math on boolean arrays is not supported")
    res = out if out is not None else ndarray.empty(a.shape, a.dtype)
    if res.size:
        L.call("pdn_ew_unary", op, _code(a.dtype), a.ptr, res.ptr, len(a.shape), _arr(a.shape), _arr(a.estrides), _arr(res.estrides))
    return res


def ternary(op, a: ndarray, b: ndarray, c: ndarray) -> ndarray:
    """Fused three-operand elementwise op (see PDN_T_* in include/pdn_b200.h); operands broadcast."""
    dt = np.result_type(a.dtype, b.dtype, c.dtype)
    if dt not in _FLOATS:
        dt = np.dtype(np.float64)
    a, b, c = (x if x.dtype == dt else x.astype(dt) for x in (a, b, c))
    shape = np.broadcast_shapes(a.shape, b.shape, c.shape)
    av, bv, cv = a.broadcast_to(shape), b.broadcast_to(shape), c.broadcast_to(shape)
    res = ndarray.empty(shape, dt)
    if res.size:
        L.call("pdn_ew_ternary", op, _code(dt), av.ptr, bv.ptr, cv.ptr, res.ptr, len(shape), _arr(shape), _arr(av.estrides),
               _arr(bv.estrides), _arr(cv.estrides), _arr(res.estrides))
    return res


def _norm_axes(axis, ndim):
    if axis is None:
        return tuple(range(ndim))
    if isinstance(axis, numbers.Integral):
        axis = (int(axis), )
    out = []
    for ax in axis:
        ax = int(ax)
        if ax < -ndim or ax >= ndim:
            raise np.exceptions.AxisError(ax, ndim)
        ax %= ndim
        if ax in out:
            raise ValueError("duplicate value in 'axis'")
        out.append(ax)
    return tuple(out)


def _reduce(op, a: ndarray, axis, keepdims) -> ndarray:
    arg = op in (L.R_ARGMAX, L.R_ARGMIN)
    if arg and axis is not None and not isinstance(axis, numbers.Integral):
        raise TypeError("argmax/argmin take a single integer axis")
    axes = _norm_axes(axis, a.ndim)
    x = a
    if x.dtype == np.bool_:
        x = x.astype(np.int64)
    if op == L.R_MEAN and x.dtype not in _FLOATS:
        x = x.astype(np.float64)
    if op == L.R_SUM and x.dtype == np.int32:
        x = x.astype(np.int64)
    if arg and axis is None and not x.is_contiguous:
        x = x.copy()  # flat index must follow C order
    mask = 0
    for ax in axes:
        mask |= 1 << ax
    kept_shape = [s for i, s in enumerate(x.shape) if i not in axes]
    if (op in (L.R_MAX, L.R_MIN) or arg) and _prod([x.shape[i] for i in axes]) == 0 and _prod(kept_shape) != 0:
        raise ValueError("zero-size array to reduction operation which has no identity")
    out_dtype = np.dtype(np.int64) if arg else x.dtype
    out = ndarray.empty(kept_shape, out_dtype)
    if out.size:
        L.call("pdn_reduce", op, _code(x.dtype), x.ptr, out.ptr, x.ndim, _arr(x.shape), _arr(x.estrides), mask)
    if keepdims:
        full = [1 if i in axes else s for i, s in enumerate(x.shape)]
        out = out.reshape(full)
    return out


# ---------------------------------------------------------------------- matmul -------------------
def gemm_into(out: ndarray | None, a: ndarray, b: ndarray, bias: ndarray | None = None, accumulate=False, prec=0) -> ndarray:
    """out = a @ b (+ bias) for >=2-D operands with NumPy batch broadcasting; batch dims collapsed to <= 3."""
    assert a.ndim >= 2 and b.ndim >= 2
    M, K = a.shape[-2:]
    K2, N = b.shape[-2:]
    if K != K2:
        raise ValueError(f"matmul: Input operand 1 has a mismatch in its core dimension 0 (size {K2} is different from {K})")
    if a.ndim == 2 and b.ndim == 2 and a.dtype == _F32 and b.dtype == _F32 and (bias is None or bias.dtype == _F32):
        # the common case (every Linear layer, every 2-D matmul and its two gradient products): no batch logic, no promotion
        if out is None:
            out = ndarray.empty((M, N), _F32)
        else:
            assert out.shape == (M, N) and out.dtype == _F32 and out.estrides[-1] == 1
        if out.size == 0:
            return out
        if bias is not None:
            assert bias.shape == (N, ) and bias.is_contiguous
        out.buf.version += 1
        L.call("pdn_gemm_cached", 0, a.ptr, b.ptr, out.ptr, M, N, K, a.estrides[0], a.estrides[1], b.estrides[0], b.estrides[1], out.estrides[0],
               _ONE3, _ZERO3, _ZERO3, _ZERO3, bias.ptr if bias is not None else None, 1 if accumulate else 0, prec, a.buf.version, b.buf.version)
        return out
    batch = np.broadcast_shapes(a.shape[:-2], b.shape[:-2])
    dt = np.result_type(a.dtype, b.dtype)
    if dt not in _FLOATS:
        dt = np.dtype(np.float64)
    if a.dtype != dt:
        a = a.astype(dt)
    if b.dtype != dt:
        b = b.astype(dt)
    # (…, M, K) @ (K, N) with mergeable leading dims of `a` is one flat GEMM over M' = prod(batch)*M
    if out is None and b.ndim == 2 and a.ndim > 2:
        st = _reshape_strides(a.shape, a.estrides, (_prod(a.shape[:-1]), K))
        if st is not None:
            flat = a._view((_prod(a.shape[:-1]), K), st)
            r = gemm_into(None, flat, b, bias, False, prec)
            return r.reshape(tuple(a.shape[:-1]) + (N, ))
    if out is None:
        out = ndarray.empty(tuple(batch) + (M, N), dt)
    else:
        assert out.shape == tuple(batch) + (M, N) and out.dtype == dt and out.estrides[-1] == 1
    if out.size == 0:
        return out
    ab = a.broadcast_to(tuple(batch) + (M, K))
    bb = b.broadcast_to(tuple(batch) + (K, N))
    # collapse batch dims jointly for a, b, out
    bs, sa, sb, sc = [], [], [], []
    for i, n in enumerate(batch):
        if n == 1:
            continue
        ea, eb, ec = ab.estrides[i], bb.estrides[i], out.estrides[i]
        if bs and sa[-1] == n * ea and sb[-1] == n * eb and sc[-1] == n * ec:
            bs[-1] *= n
            sa[-1], sb[-1], sc[-1] = ea, eb, ec
        else:
            bs.append(n); sa.append(ea); sb.append(eb); sc.append(ec)
    if len(bs) > 3:  # rare: materialise the operands contiguously and retry
        ab = ab.copy()
        bb = bb.copy()
        bs, sa, sb, sc = [_prod(batch)], [M * K], [K * N], [out.estrides[len(batch) - 1] if out.is_contiguous else None]
        if sc[0] is None:
            raise NotImplementedError("matmul: non-collapsible batch layout of the output")
        sc = [M * N]
    while len(bs) < 3:
        bs.insert(0, 1); sa.insert(0, 0); sb.insert(0, 0); sc.insert(0, 0)
    if bias is not None:
        assert bias.shape == (N, ) and bias.is_contiguous
        if bias.dtype != dt:
            bias = bias.astype(dt)
    out.buf.version += 1
    # the operands' write counters let the library keep their tensor-core operand planes across the products of one training step
    L.call("pdn_gemm_cached", _code(dt), ab.ptr, bb.ptr, out.ptr, M, N, K, ab.estrides[-2], ab.estrides[-1], bb.estrides[-2],
           bb.estrides[-1], out.estrides[-2], _arr(bs), _arr(sa), _arr(sb), _arr(sc), bias.ptr if bias is not None else None,
           1 if accumulate else 0, prec, ab.buf.version, bb.buf.version)
    return out


def matmul(a, b) -> ndarray:
    a, b = _to_dev(a), _to_dev(b)
    if a.ndim == 0 or b.ndim == 0:
        raise ValueError("matmul: Input operand does not have enough dimensions")
    exp_a, exp_b = a.ndim == 1, b.ndim == 1
    if exp_a:
        a = a._view((1, ) + a.shape, (0, ) + a.estrides)
    if exp_b:
        b = b._view(b.shape + (1, ), b.estrides + (0, ))
    r = gemm_into(None, a, b)
    if exp_a:
        r = r._view(r.shape[:-2] + r.shape[-1:], r.estrides[:-2] + r.estrides[-1:])
    if exp_b:
        r = r._view(r.shape[:-1], r.estrides[:-1])
    return r
