"""Tensor, the autograd tape and the eager operators.

Public surface = reference pydynet/core/tensor.py (Tensor :30-413, 27 operators :535-1019); results are the
reference's, the machinery is not:

* no global creation-ordered node list (reference tensor.py:9-27, O(N^2) bookkeeping): every result carries a
  ``_Node`` (backward closure + input edges + creation sequence number) and ``backward()`` sorts the *ancestors*
  of the root by that number;
* one ``backward`` call per node producing all input grads (the reference calls ``grad_fn`` per edge), so e.g. the
  two GEMMs of a matmul node or the three outputs of a fused attention node share work;
* gradients of interior nodes are transient and allocated on first contribution (the reference eagerly zero-fills
  a ``grad`` buffer per node, tensor.py:89-90); only leaves keep ``.grad`` across calls;
* on a cuda device every array expression below runs in libpdn_b200.so kernels (``Device.xp`` is
  pydynet_b200.backend); on the cpu device ``xp`` is NumPy exactly like the reference.
"""
from __future__ import annotations

import itertools
import numbers

import numpy as np

from ..autograd import is_grad_enable, no_grad
from ..cuda import Device
from ..backend.array import ndarray as _DevArray

_seq = itertools.count()


def _is_dev(a) -> bool:
    return not isinstance(a, (np.ndarray, np.generic))


def _owns_memory(g, upstream) -> bool:
    """True if array ``g`` is a fresh result that nothing else aliases (safe to mutate in place / keep)."""
    if g is upstream:
        return False
    if isinstance(g, np.ndarray):
        return g.base is None and g.flags.writeable
    if isinstance(g, np.generic):
        return False
    if upstream is not None and _is_dev(upstream) and g.buf is upstream.buf:
        return False
    return g.is_contiguous


class _Node:
    """Tape entry of one operator result."""
    __slots__ = ("inputs", "backward", "seq", "name")

    def __init__(self, inputs, backward, name):
        self.inputs = inputs  # tuple[Tensor]
        self.backward = backward  # grad(array) -> tuple[array | None] aligned with inputs
        self.seq = next(_seq)
        self.name = name


_LEAF_READY_HOOK = [None]  # callable(leaf | None) installed by distributed.DataParallel(overlap=True); None = sweep finished


class Tensor:
    """Differentiable array (reference tensor.py:30-413)."""
    _pdn_hint = None  # set on logits produced by an inference plan (nn/_plans.py): memoised argmax over the vocabulary

    def __init__(self, data, dtype=None, copy=True, device=None, requires_grad: bool = False) -> None:
        if isinstance(data, Tensor):
            raise ValueError('Tensor assignment with another tensor is forbidden.')
        self.device = Device(device)
        xp = self.device.xp
        with self.device:
            if xp is np:
                if isinstance(data, _DevArray):
                    data = data.get()  # device array handed to a cpu tensor
                self.data = np.array(data, dtype=dtype, copy=copy)
            else:
                self.data = xp.array(data, dtype=dtype, copy=bool(copy))
        self._node = None
        self._grad = None
        self._pinned_grad = False  # grad storage is a view into a flat bucket (optim / data-parallel)
        self._grad_stale = False  # pinned storage holds garbage/old values: next accumulation overwrites
        self.requires_grad = bool(is_grad_enable() and requires_grad)
        if self.requires_grad:
            if not np.issubdtype(self.data.dtype, np.floating):
                raise TypeError("Only Tensors of floating point dtype can require gradients!")
            # reference quirk (tensor.py:90): the grad buffer takes the *constructor's* dtype argument, so a leaf built
            # without dtype accumulates in float64 whatever its data dtype is.
            self._grad_dtype = np.dtype(dtype) if dtype is not None else np.dtype(np.float64)

    # ------------------------------------------------------------------ autograd state ----------
    @property
    def grad(self):
        if not self.requires_grad:
            return None
        if self._grad is None:
            with self.device:
                self._grad = self.xp.zeros(self.shape, dtype=self._grad_dtype)
        elif self._grad_stale:
            with self.device:
                self._grad[...] = 0.
            self._grad_stale = False
        return self._grad

    @grad.setter
    def grad(self, value):
        if self._pinned_grad and self._grad is not None and value is not None and value is not self._grad:
            # the storage is a view into a flat bucket (fused Adam / data-parallel all-reduce read THAT memory): write through
            with self.device:
                self._grad[...] = value
        else:
            self._grad = value
        self._grad_stale = False

    @property
    def last(self):
        """Upstream tensors of this node (reference attribute name)."""
        return list(self._node.inputs) if self._node is not None else []

    @property
    def is_leaf(self) -> bool:
        return not self.requires_grad or self._node is None

    @property
    def xp(self):
        return self.device.xp

    # ------------------------------------------------------------------ metadata ----------------
    @property
    def shape(self):
        return self.data.shape

    @property
    def ndim(self):
        return self.data.ndim

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def size(self):
        return self.data.size

    @property
    def strides(self):
        return self.data.strides

    @property
    def T(self):
        return self.transpose()

    def __repr__(self) -> str:
        if self._node is not None:
            return "Tensor({}, op={})".format(self.data, self._node.name)
        return "Tensor({}, requires_grad={}".format(self.data, self.requires_grad) + (", device={}".format(
            self.device) if self.device.device != "cpu" else "") + ")"

    def __len__(self) -> int:
        return len(self.data)

    # ------------------------------------------------------------------ conversions --------------
    def astype(self, new_type) -> "Tensor":
        assert not self.requires_grad
        with self.device:
            return Tensor(self.data.astype(new_type), new_type, copy=None, device=self.device)

    def numpy(self):
        if self.device.is_cuda:
            with self.device:
                return self.data.get()
        return self.data.copy()

    def item(self):
        with self.device:
            return self.data.item()

    def to(self, device) -> "Tensor":
        device = Device(device)
        if self.device != device:
            old = self.device
            grad = self._grad
            if grad is not None and self._grad_stale:
                grad = None
            self.device = device
            self.data = _move(self.data, old, device)
            if self.requires_grad:
                self._grad = _move(grad, old, device) if grad is not None else None
                self._pinned_grad = False
                self._grad_stale = False
        return self

    def cpu(self):
        return self.to("cpu")

    def cuda(self, id: int = 0):
        return self.to(f"cuda:{id}")

    # ------------------------------------------------------------------ methods → operators ------
    def reshape(self, *new_shape): return reshape(self, new_shape)
    def transpose(self, *axes): return transpose(self, axes if len(axes) != 0 else None)
    def swapaxes(self, axis1, axis2): return swapaxes(self, axis1, axis2)
    def max(self, axis=None, keepdims=False): return max(self, axis, keepdims)
    def min(self, axis=None, keepdims=False): return min(self, axis, keepdims)
    def mean(self, axis=None, keepdims=False): return mean(self, axis, keepdims)
    def sum(self, axis=None, keepdims=False): return sum(self, axis, keepdims)
    def argmax(self, axis=None, keepdims=False): return argmax(self, axis, keepdims)
    def argmin(self, axis=None, keepdims=False): return argmin(self, axis, keepdims)
    def __add__(self, x): return add(self, x)
    def __radd__(self, x): return add(x, self)
    def __sub__(self, x): return sub(self, x)
    def __rsub__(self, x): return sub(x, self)
    def __mul__(self, x): return mul(self, x)
    def __rmul__(self, x): return mul(x, self)
    def __matmul__(self, x): return matmul(self, x)
    def __rmatmul__(self, x): return matmul(x, self)
    def __truediv__(self, x): return div(self, x)
    def __rtruediv__(self, x): return div(x, self)
    def __pow__(self, x): return pow(self, x)
    def __rpow__(self, x): return pow(x, self)
    def __pos__(self): return 1 * self
    def __neg__(self): return -1 * self
    def __abs__(self): return abs(self)
    def __getitem__(self, key): return _get_slice(self, key)

    # in-place family (reference tensor.py:266-294): forbidden on grad-requiring tensors while grad mode is on
    def _inplace(self, func_name, *others):
        if self.requires_grad and is_grad_enable():
            raise ValueError("In-place operation is forbidden in node requires grad.")
        others = tuple(o.data if isinstance(o, Tensor) else o for o in others)
        with self.device:
            r = getattr(self.data, func_name)(*others)
            if func_name != "__setitem__" and r is not NotImplemented and r is not None:
                self.data = r
        return self

    def __setitem__(self, key, value):
        if isinstance(key, tuple):
            key = tuple(k.data if isinstance(k, Tensor) else k for k in key)
        elif isinstance(key, Tensor):
            key = key.data
        return self._inplace("__setitem__", key, value)

    def __iadd__(self, other): return self._inplace("__iadd__", other)
    def __isub__(self, other): return self._inplace("__isub__", other)
    def __imul__(self, other): return self._inplace("__imul__", other)
    def __itruediv__(self, other): return self._inplace("__itruediv__", other)
    def __imatmul__(self, other): return self._inplace("__imatmul__", other)

    # comparisons return non-differentiable bool tensors (reference tensor.py:296-325); no __eq__ on purpose
    def _compare(self, other, fn):
        with self.device:
            return Tensor(fn(self.data, other.data if isinstance(other, Tensor) else other), np.bool_, None, self.device, False)

    def eq(self, other): return self._compare(other, lambda x, y: x == y)
    def ne(self, other): return self._compare(other, lambda x, y: x != y)
    def __lt__(self, other): return self._compare(other, lambda x, y: x < y)
    def __le__(self, other): return self._compare(other, lambda x, y: x <= y)
    def __gt__(self, other): return self._compare(other, lambda x, y: x > y)
    def __ge__(self, other): return self._compare(other, lambda x, y: x >= y)

    # ------------------------------------------------------------------ backward ----------------
    def zero_grad(self):
        if self._pinned_grad:
            self._grad_stale = True  # storage stays, next accumulation overwrites (no memset launch)
        else:
            self._grad = None  # materialised lazily as zeros if read

    def _accumulate(self, g, owned: bool):
        """leaf.grad += g (reference tensor.py:371) without touching memory that is known to be zero."""
        xp = self.xp
        if self._grad is not None and not self._grad_stale:
            self._grad += g
            return
        if self._pinned_grad:
            self._grad[...] = g
            self._grad_stale = False
            return
        if xp is np:
            self._grad = np.array(g, dtype=self._grad_dtype, order="C", copy=None if owned else True)
            return
        if g.dtype != self._grad_dtype:
            g = g.astype(self._grad_dtype)
        elif not owned or not g.is_contiguous:
            g = g.copy()
        self._grad = g

    def backward(self, retain_graph: bool = False):
        """Reverse sweep from a scalar (reference tensor.py:327-375). Errors match the reference's: ValueError when the
        tensor is not part of a graph or is not a scalar."""
        if not self.requires_grad or (self._node is None and getattr(self, "_freed", False)):
            raise ValueError("Auto-grad is failed because current node is not in graph.")
        if self.size > 1:
            raise ValueError("backward should be called only on a scalar.")
        from .. import cuda as _cuda
        if _cuda.NVTX and self.device.is_cuda:
            with _cuda.nvtx_range("backward"):
                _cuda.NVTX = False  # one range per sweep, not one per nested call
                try:
                    return self.backward(retain_graph)
                finally:
                    _cuda.NVTX = True
        with self.device:
            seed = self.xp.ones(self.shape, dtype=self.dtype)
            if self._node is None:  # a leaf: d self / d self
                self._accumulate(seed, True)
                return
            # ancestors of the root that carry a tape entry, newest first
            order, seen, stack = [], {id(self)}, [self]
            while stack:
                t = stack.pop()
                order.append(t)
                for i in t._node.inputs:
                    if i.requires_grad and i._node is not None and id(i) not in seen:
                        seen.add(id(i))
                        stack.append(i)
            order.sort(key=lambda t: -t._node.seq)
            # data-parallel overlap (distributed.DataParallel): a leaf whose gradient lives in the flat bucket is reported the moment
            # its LAST consumer in this sweep has run, so that the all-reduce of a finished bucket range can start on the side
            # stream while the rest of the backward pass is still computing
            hook, uses = _LEAF_READY_HOOK[0], None
            if hook is not None:
                uses = {}
                for t in order:
                    for i in t._node.inputs:
                        if i._pinned_grad and i._node is None and i.requires_grad:
                            uses[id(i)] = uses.get(id(i), 0) + 1
            pending = {id(self): [seed, True]}
            for t in order:
                slot = pending.pop(id(t), None)
                node = t._node
                if uses:
                    leaves = [i for i in node.inputs if id(i) in uses]
                if slot is not None:
                    gout = slot[0]
                    if retain_graph:
                        t._grad = gout
                    grads = node.backward(gout)
                    for inp, g in zip(node.inputs, grads):
                        if g is None or not inp.requires_grad:
                            continue
                        owned = _owns_memory(g, gout)
                        if inp._node is None:
                            if not getattr(inp, "_freed", False):
                                inp._accumulate(g, owned)
                            continue
                        s = pending.get(id(inp))
                        if s is None:
                            pending[id(inp)] = [g, owned]
                        elif s[1]:
                            s[0] += g
                        else:
                            s[0] = s[0] + g
                            s[1] = True
                if uses:
                    for i in leaves:
                        uses[id(i)] -= 1
                        if uses[id(i)] == 0:
                            hook(i)
                if not retain_graph:
                    t._node = None
                    t._freed = True
            if hook is not None:
                hook(None)  # sweep finished

    # kept for API compatibility with user code that wires edges by hand
    def _build_edge(self, node: "Tensor"):
        if node._node is None:
            node._node = _Node((self, ), lambda g: (None, ), type(node).__name__)
        elif self not in node._node.inputs:
            node._node.inputs = tuple(node._node.inputs) + (self, )


def _is_contig(a) -> bool:
    if isinstance(a, np.ndarray):
        return a.flags.c_contiguous
    return a.is_contiguous


def _move(arr, old: Device, new: Device):
    """Array migration between devices: host<->device copies go through pdn_memcpy_h2d/d2h."""
    if old.is_cuda:
        with old:
            host = arr.get()
    else:
        host = arr
    if new.is_cuda:
        with new:
            return new.xp.array(host)
    return host


def _wrap(x, like: Tensor | None = None) -> Tensor:
    if isinstance(x, Tensor):
        return x
    if like is not None:
        return Tensor(x, dtype=like.dtype, device=like.device)
    return Tensor(x)


def _result(data, device: Device, inputs, backward, name) -> Tensor:
    """Wraps an operator result; records the tape entry when any input requires grad and grad mode is on."""
    out = Tensor.__new__(Tensor)
    out.device = device
    out.data = data
    out._grad = None
    out._pinned_grad = False
    out._grad_stale = False
    out._node = None
    rg = False
    if is_grad_enable():
        for i in inputs:
            if i.requires_grad:
                rg = True
                break
    if rg and not np.issubdtype(data.dtype, np.floating):
        raise TypeError("Only Tensors of floating point dtype can require gradients!")
    out.requires_grad = rg
    if rg:
        out._grad_dtype = data.dtype
        out._node = _Node(tuple(inputs), backward, name)
    return out


_SCALARS = {}  # (device, dtype, value bits) -> constant Tensor already resident on that cuda device


def _wrap_scalar(value, like):
    """``Tensor(value, dtype=like.dtype, device=like.device)`` (reference tensor.py:488-493). On a cuda device a Python / NumPy
    SCALAR is uploaded once and the constant Tensor is reused: the reference's idioms wrap one on every call
    (``relu = maximum(0., x)``, ``x * 2``, ``1 - z``), and every fresh wrap costs an allocation plus a SYNCHRONOUS 4-byte
    host-to-device copy — a stream drain per ReLU in a launch-bound training step."""
    if like.device.is_cuda and isinstance(value, (int, float, np.floating, np.integer)) and not isinstance(value, bool):
        from ..cuda import is_capturing
        v = np.asarray(value, dtype=like.dtype)
        key = (like.device, v.dtype.str, v.tobytes())
        t = _SCALARS.get(key)
        if t is not None:
            return t  # (also while a CUDA graph is being recorded: the constant was allocated outside the recording and is never freed)
        if is_capturing():
            # a constant first seen INSIDE a recording: a host-to-device copy cannot be recorded, a fill kernel can; its memory belongs to
            # the graph, so it is not cached
            with like.device:
                return _result(like.device.xp.full((), v.item(), dtype=v.dtype), like.device, (), None, "const")
        if len(_SCALARS) < 4096:  # never evicted: recorded graphs may reference the cached buffers
            t = _SCALARS[key] = Tensor(value, dtype=like.dtype, device=like.device)
            return t
    return Tensor(value, dtype=like.dtype, device=like.device)


def _binary_prepare(x, y):
    """Scalar/array operands take the dtype and device of the Tensor operand (reference tensor.py:488-494)."""
    if not isinstance(x, Tensor) and isinstance(y, Tensor):
        x = _wrap_scalar(x, y)
    elif isinstance(x, Tensor) and not isinstance(y, Tensor):
        y = _wrap_scalar(y, x)
    elif not (isinstance(x, Tensor) and isinstance(y, Tensor)):
        x, y = Tensor(x), Tensor(y)
    assert x.device == y.device
    return x, y


def _sum_to(g, shape):
    """Un-broadcast: reduce ``g`` to ``shape`` (the engine-side logic of reference tensor.py:360-370, done per operand)."""
    if g.shape == tuple(shape):
        return g
    extra = g.ndim - len(shape)
    axes = tuple(range(extra)) + tuple(i + extra for i, s in enumerate(shape) if s == 1 and g.shape[i + extra] != 1)
    if axes:
        g = g.sum(axis=axes, keepdims=True)
    return g.reshape(shape)


def _ext(xp):
    """Fused helpers: device kernels for cuda, NumPy expressions for the cpu device."""
    if xp is np:
        from . import _host_ext
        return _host_ext
    return xp.ext


# ====================================================================== arithmetic =================
def add(x, y) -> Tensor:
    x, y = _binary_prepare(x, y)
    with x.device:
        data = x.data + y.data

    def backward(g):
        return (_sum_to(g, x.shape) if x.requires_grad else None, _sum_to(g, y.shape) if y.requires_grad else None)

    return _result(data, x.device, (x, y), backward, "add")


def sub(x, y) -> Tensor:
    x, y = _binary_prepare(x, y)
    with x.device:
        data = x.data - y.data

    def backward(g):
        return (_sum_to(g, x.shape) if x.requires_grad else None, _sum_to(-g, y.shape) if y.requires_grad else None)

    return _result(data, x.device, (x, y), backward, "sub")


def mul(x, y) -> Tensor:
    x, y = _binary_prepare(x, y)
    with x.device:
        data = x.data * y.data

    def backward(g):
        return (_sum_to(y.data * g, x.shape) if x.requires_grad else None,
                _sum_to(x.data * g, y.shape) if y.requires_grad else None)

    return _result(data, x.device, (x, y), backward, "mul")


def div(x, y) -> Tensor:
    x, y = _binary_prepare(x, y)
    with x.device:
        data = x.data / y.data

    def backward(g):
        t = g / y.data
        return (_sum_to(t, x.shape) if x.requires_grad else None, _sum_to(-data * t, y.shape) if y.requires_grad else None)

    return _result(data, x.device, (x, y), backward, "div")


def pow(x, y) -> Tensor:
    x, y = _binary_prepare(x, y)
    with x.device:
        data = x.data**y.data

    def backward(g):
        gx = _sum_to((data * y.data / x.data) * g, x.shape) if x.requires_grad else None
        gy = _sum_to(data * x.xp.log(x.data) * g, y.shape) if y.requires_grad else None
        return gx, gy

    return _result(data, x.device, (x, y), backward, "pow")


def maximum(x, y) -> Tensor:
    """Ties send the gradient to both operands (reference tensor.py:808-815), hence relu'(0) = 1."""
    x, y = _binary_prepare(x, y)
    xp = x.xp
    with x.device:
        data = xp.maximum(x.data, y.data)

    def backward(g):
        e = _ext(xp)
        return (_sum_to(e.eq_mul(data, x.data, g), x.shape) if x.requires_grad else None,
                _sum_to(e.eq_mul(data, y.data, g), y.shape) if y.requires_grad else None)

    return _result(data, x.device, (x, y), backward, "maximum")


def minimum(x, y) -> Tensor:
    x, y = _binary_prepare(x, y)
    xp = x.xp
    with x.device:
        data = xp.minimum(x.data, y.data)

    def backward(g):
        e = _ext(xp)
        return (_sum_to(e.eq_mul(data, x.data, g), x.shape) if x.requires_grad else None,
                _sum_to(e.eq_mul(data, y.data, g), y.shape) if y.requires_grad else None)

    return _result(data, x.device, (x, y), backward, "minimum")


def matmul(x, y) -> Tensor:
    """NumPy ``@`` semantics incl. 1-D promotion and broadcast batch dims (reference tensor.py:643-676)."""
    x, y = _binary_prepare(x, y)
    xp = x.xp
    with x.device:
        data = x.data @ y.data
    exp_a, exp_b = x.ndim < 2, y.ndim < 2

    def backward(g):
        # restore the (..., M, N) layout of the product for promoted 1-D operands
        if exp_b:
            g = xp.expand_dims(g, -1)
        if exp_a:
            g = xp.expand_dims(g, -2)
        a = x.data.reshape(1, -1) if exp_a else x.data  # (..., M, K)
        b = y.data.reshape(-1, 1) if exp_b else y.data  # (..., K, N)
        gx = gy = None
        if x.requires_grad:
            gx = g @ b.swapaxes(-1, -2)
            if exp_a:
                gx = gx[..., 0, :]
            gx = _sum_to(gx, x.shape)
        if y.requires_grad:
            gy = _ext(xp).matmul_dB(a, g, b.shape)
            if exp_b:
                gy = gy[..., 0]
            gy = _sum_to(gy, y.shape)
        return gx, gy

    return _result(data, x.device, (x, y), backward, "matmul")


# ====================================================================== unary ======================
def _unary(name, x, fwd, bwd) -> Tensor:
    x = _wrap(x)
    with x.device:
        data = fwd(x.xp, x.data)
    return _result(data, x.device, (x, ), lambda g: (bwd(x.xp, x.data, data, g), ), name)


def abs(x) -> Tensor:
    # the reference's abs.grad_fn is broken (passes a Tensor to xp.sign, tensor.py:692); we define the obvious grad.
    return _unary("abs", x, lambda xp, a: xp.abs(a), lambda xp, a, out, g: g * xp.sign(a))


def exp(x) -> Tensor:
    return _unary("exp", x, lambda xp, a: xp.exp(a), lambda xp, a, out, g: out * g)


def log(x) -> Tensor:
    return _unary("log", x, lambda xp, a: xp.log(a), lambda xp, a, out, g: g / a)


def sign(x) -> Tensor:
    return _unary("sign", x, lambda xp, a: xp.sign(a), lambda xp, a, out, g: xp.zeros(out.shape, dtype=out.dtype))


def sigmoid(x) -> Tensor:
    """Overflow-safe piecewise form of the reference (tensor.py:996-1005)."""
    return _unary("sigmoid", x, lambda xp, a: _ext(xp).sigmoid(a), lambda xp, a, out, g: _ext(xp).sigmoid_grad(out, g))


def tanh(x) -> Tensor:
    return _unary("tanh", x, lambda xp, a: _ext(xp).tanh(a), lambda xp, a, out, g: _ext(xp).tanh_grad(out, g))


# ====================================================================== reductions =================
def _norm_axis(axis, ndim):
    if axis is None:
        return None
    if isinstance(axis, numbers.Integral):
        return (int(axis) % ndim if ndim else 0, )
    return tuple(int(a) % ndim for a in axis)


def _expand_reduced(xp, g, axis, keepdims):
    if axis is None or keepdims:
        return g
    return xp.expand_dims(g, axis)


def sum(x, axis=None, keepdims=False) -> Tensor:
    x = _wrap(x)
    xp = x.xp
    with x.device:
        data = x.data.sum(axis=axis, keepdims=keepdims)

    def backward(g):
        return (xp.broadcast_to(_expand_reduced(xp, g, axis, keepdims), x.shape), )

    return _result(xp.asarray(data), x.device, (x, ), backward, "sum")


def mean(x, axis=None, keepdims=False) -> Tensor:
    x = _wrap(x)
    xp = x.xp
    with x.device:
        data = xp.asarray(x.data.mean(axis=axis, keepdims=keepdims))

    def backward(g):
        scale = data.size / x.size
        return (xp.broadcast_to(_expand_reduced(xp, g * scale, axis, keepdims), x.shape), )

    return _result(data, x.device, (x, ), backward, "mean")


def _minmax(name, x, axis, keepdims) -> Tensor:
    x = _wrap(x)
    xp = x.xp
    with x.device:
        data = xp.asarray(getattr(x.data, name)(axis=axis, keepdims=keepdims))

    def backward(g):
        # every element equal to the extremum receives the full gradient (reference tensor.py:741-747)
        full = _expand_reduced(xp, data, axis, keepdims)
        return (_ext(xp).eq_mul(full, x.data, _expand_reduced(xp, g, axis, keepdims)), )

    return _result(data, x.device, (x, ), backward, name)


def max(x, axis=None, keepdims=False) -> Tensor:
    return _minmax("max", x, axis, keepdims)


def min(x, axis=None, keepdims=False) -> Tensor:
    return _minmax("min", x, axis, keepdims)


def _arg(name, x, axis, keepdims) -> Tensor:
    x = _wrap(x)
    with x.device:
        data = x.xp.asarray(getattr(x.data, name)(axis=axis, keepdims=keepdims))
    # int64 result: raises TypeError on a grad-requiring input like the reference (tensor.py:85-88)
    return _result(data, x.device, (x, ), None, name)


def argmax(x, axis=None, keepdims=False) -> Tensor:
    if isinstance(x, Tensor) and x._pdn_hint is not None:
        from ..nn import _plans
        memo = _plans.hint_argmax(x, axis, keepdims)
        if memo is not None:
            return memo
    return _arg("argmax", x, axis, keepdims)


def argmin(x, axis=None, keepdims=False) -> Tensor:
    return _arg("argmin", x, axis, keepdims)


# ====================================================================== shape ops (views) ==========
def reshape(x, new_shape) -> Tensor:
    x = _wrap(x)
    if len(new_shape) == 1 and isinstance(new_shape[0], (tuple, list)):
        new_shape = tuple(new_shape[0])
    with x.device:
        data = x.data.reshape(new_shape)
    return _result(data, x.device, (x, ), lambda g: (g.reshape(x.shape), ), "reshape")


def transpose(x, axes=None) -> Tensor:
    x = _wrap(x)
    if axes is not None and len(axes) == 1 and isinstance(axes[0], (tuple, list)):
        axes = tuple(axes[0])
    with x.device:
        data = x.data.transpose(axes) if axes is not None else x.data.transpose()

    def backward(g):
        if axes is None:
            return (g.transpose(), )
        return (g.transpose(tuple(int(i) for i in np.argsort([a % x.ndim for a in axes]))), )

    return _result(data, x.device, (x, ), backward, "transpose")


def swapaxes(x, axis1, axis2) -> Tensor:
    x = _wrap(x)
    with x.device:
        data = x.data.swapaxes(axis1, axis2)
    return _result(data, x.device, (x, ), lambda g: (g.swapaxes(axis1, axis2), ), "swapaxes")


def _get_slice(x, key) -> Tensor:
    """x[key]; backward is zeros + *assignment* so duplicate indices are last-write-wins, not accumulated
    (reference tensor.py:934-940)."""
    x = _wrap(x)
    if x._pdn_hint is not None:  # logits of an inference plan: ``logits[:, -1, :]`` is a ready-made view carrying the argmax memo
        from ..nn import _plans
        out = _plans.hint_getitem(x, key)
        if out is not None:
            return out
    if isinstance(key, tuple):
        key = tuple(k.data if isinstance(k, Tensor) else k for k in key)
    elif isinstance(key, Tensor):
        key = key.data
    xp = x.xp
    with x.device:
        data = xp.asarray(x.data[key])

    def backward(g):
        full = xp.zeros(x.shape, dtype=x.dtype)
        full[key] = g
        return (full, )

    return _result(data, x.device, (x, ), backward, "_get_slice")


def concat(tensors, axis=0) -> Tensor:
    tensors = list(tensors)
    for t in tensors:
        assert isinstance(t, Tensor), "Concatenate elements in 'tensors' must be 'Tensor'"
        assert t.device == tensors[0].device
    dev = tensors[0].device
    with dev:
        data = dev.xp.concatenate([t.data for t in tensors], axis=axis)
    bounds = [0]
    for t in tensors:
        bounds.append(bounds[-1] + t.shape[axis])

    def backward(g):
        outs = []
        for i, t in enumerate(tensors):
            if not t.requires_grad:
                outs.append(None)
                continue
            slc = [slice(None)] * g.ndim
            slc[axis] = slice(bounds[i], bounds[i + 1])
            outs.append(g[tuple(slc)])
        return tuple(outs)

    return _result(data, dev, tensors, backward, "concat")


# ====================================================================== user-extensible protocol ===
class _UnaryOperator(Tensor):
    """Reference operator protocol (tensor.py:416-467): subclass, implement ``forward_(x) -> array`` and
    ``grad_fn(x, grad) -> array``. Kept so user-defined operators written for the reference keep working."""

    def __init__(self, x) -> None:
        if not isinstance(x, Tensor):
            x = Tensor(x)
        self.device = x.device
        with self.device:
            data = self.forward_(x)
        Tensor.__init__(self, data, dtype=data.dtype, copy=None, device=x.device)
        self.data = data
        if is_grad_enable() and x.requires_grad:
            self.requires_grad = True
            self._grad_dtype = data.dtype
            self._node = _Node((x, ), lambda g: (self.grad_fn(x, g), ), type(self).__name__)

    def forward_(self, x):
        raise NotImplementedError

    def grad_fn(self, x, grad):
        raise NotImplementedError


class _BinaryOperator(Tensor):
    """Reference operator protocol for two operands (tensor.py:470-533)."""

    def __init__(self, x, y) -> None:
        x, y = _binary_prepare(x, y)
        self.device = x.device
        with self.device:
            data = self.forward_(x, y)
        Tensor.__init__(self, data, dtype=data.dtype, copy=None, device=x.device)
        self.data = data
        if is_grad_enable() and (x.requires_grad or y.requires_grad):
            self.requires_grad = True
            self._grad_dtype = data.dtype

            def backward(g):
                return (_sum_to(self.grad_fn(x, g), x.shape) if x.requires_grad else None,
                        _sum_to(self.grad_fn(y, g), y.shape) if y.requires_grad else None)

            self._node = _Node((x, y), backward, type(self).__name__)

    def forward_(self, x, y):
        raise NotImplementedError

    def grad_fn(self, x, grad):
        raise NotImplementedError
