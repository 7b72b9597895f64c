"""Functional layer (surface of reference nn/functional.py:7-381).

Every function has ONE definition of its result — the reference's chain of eager operators — and two ways of
computing it: the chain itself (any device, any dtype; this is what runs on the cpu device) and, for fp32 tensors on a
cuda device, a single fused autograd node backed by a hand-written kernel (nn/_fused.py → libpdn_b200.so).  The fused
nodes are parity-tested against the chain and the oracle; ``PDN_FUSED=0`` forces the chain on cuda for debugging.
"""
import numpy as np

from ..core import tensor, function
from ..core.tensor import Tensor, _result
from ..core.function import unsqueeze  # noqa: F401  (the reference's functional.py re-exports it, functional.py:4)
from ..autograd import no_grad
from . import _fused


def linear(x: Tensor, weight: Tensor, bias: Tensor):
    if _fused.usable(x, weight, bias, op='linear'):
        return _fused.linear(x, weight, bias)
    affine = x @ weight
    if bias is not None:
        affine = affine + bias
    return affine


def embedding(x: Tensor, weight: Tensor, padding_idx: int):
    query = weight[x]
    if padding_idx is not None:
        with no_grad():
            mask = function.unsqueeze(x.ne(padding_idx), -1)
        query = query * mask
    return query


def sigmoid(x: Tensor):
    return tensor.sigmoid(x)


def tanh(x: Tensor):
    return tensor.tanh(x)


def relu(x: Tensor):
    return tensor.maximum(0., x)


def leaky_relu(x: Tensor, alpha: float):
    return tensor.maximum(x, alpha * x)


def silu(x: Tensor):
    if _fused.usable(x, op='silu'):
        return _fused.silu(x)
    return x / (1 + tensor.exp(-x))


def _is_last_axis(x, axis):
    return axis is not None and not isinstance(axis, (tuple, list)) and x.ndim > 0 and axis % x.ndim == x.ndim - 1


def softmax(x: Tensor, axis=None):
    if _is_last_axis(x, axis) and _fused.usable(x, op='softmax'):
        return _fused.softmax(x, log=False)
    with no_grad():
        max_ = x.max(axis, keepdims=True)
    exp_ = tensor.exp(x - max_)
    return exp_ / tensor.sum(exp_, axis=axis, keepdims=True)


def log_softmax(x: Tensor, axis=None, keepdims=False):
    if _is_last_axis(x, axis) and keepdims and _fused.usable(x, op='softmax'):
        return _fused.softmax(x, log=True)
    with no_grad():
        max_ = x.max(axis, keepdims=True)
    x_sub_max = x - max_
    return x_sub_max - tensor.log(tensor.sum(tensor.exp(x_sub_max), axis=axis, keepdims=keepdims))


# ---------------------------------------------------------------------- im2col family (cpu device) ---
def _pad_nd(x: Tensor, pad: int, nsp: int) -> Tensor:
    """Zero padding of the last ``nsp`` axes as one autograd node (reference functional.py:97-110, 235-251)."""
    if pad == 0:
        return x
    xp = x.xp
    with x.device:
        data = xp.pad(x.data, [(0, 0)] * (x.ndim - nsp) + [(pad, pad)] * nsp, 'constant')
    inner = (Ellipsis, ) + (slice(pad, -pad), ) * nsp
    return _result(data, x.device, (x, ), lambda g: (g[inner], ), "pad")


def _im2col(x: Tensor, k: int, stride: int, nsp: int) -> Tensor:
    """Sliding windows copied out as (N, C, k[, k], out[, out]); backward scatter-adds them back (col2im) — reference
    functional.py:61-94, 194-232.  NumPy: strided view + copy / np.add.at.  cuda (only reached with PDN_FUSED=0 or for
    the 1-D family): integer-array gather / accumulate-scatter kernels with the same window index arithmetic."""
    xp, a = x.xp, x.data
    outs = tuple((a.shape[-nsp + i] - k) // stride + 1 for i in range(nsp))
    shape = a.shape[:2] + (k, ) * nsp + outs
    if xp is np:
        sp = a.strides[-nsp:]
        strides = a.strides[:2] + sp + tuple(s * stride for s in sp)
        col = np.lib.stride_tricks.as_strided(a, shape=shape, strides=strides).copy()

        def backward(g):
            gx = np.zeros(a.shape, dtype=g.dtype)
            gsp = gx.strides[-nsp:]
            view = np.lib.stride_tricks.as_strided(gx, shape=shape, strides=gx.strides[:2] + gsp + tuple(s * stride for s in gsp))
            np.add.at(view, (Ellipsis, ), g)
            return (gx, )
    else:
        win = np.arange(k)[:, None] + stride * np.arange(max(outs))[None, :]  # (k, out) source coordinate
        if nsp == 1:
            key = (slice(None), slice(None), win[:, :outs[0]])
        else:
            key = (slice(None), slice(None), win[:, None, :outs[0], None], win[None, :, None, :outs[1]])
        with x.device:
            col = a[key]

        def backward(g):
            gx = xp.zeros(a.shape, dtype=g.dtype)
            xp.add_at(gx, key, g)
            return (gx, )

    return _result(col, x.device, (x, ), backward, "im2col")


def conv1d(x: Tensor, kernel: Tensor, padding: int = 0, stride: int = 1):
    """1-D convolution (reference functional.py:113-140). The reference's expression ``(col @ kernel.transpose(1, 2, 0)).sum(1)``
    multiplies (k, n_out) windows by (k, O) kernels, so it is only defined when the number of output positions equals the kernel size
    (it raises otherwise) and then contracts the window POSITIONS with the kernel taps: out[n, o, j] = sum_{c,l} x_pad[n, c, j + l*stride]
    * kernel[o, c, l]. Exactly that is computed for those inputs (pinned by tests/golden/family_1d.npz); where the reference raises,
    this is the ordinary strided correlation out[n, o, l] = sum_{c,j} x_pad[n, c, l*stride + j] * kernel[o, c, j]."""
    col = _im2col(_pad_nd(x, padding, 1), kernel.shape[-1], stride, 1)  # (N, C, k, L)
    N, C, k, L = col.shape
    if L == k:
        return (col @ kernel.transpose(1, 2, 0)).sum(1).swapaxes(1, 2)
    out = col.transpose(0, 3, 1, 2).reshape(N * L, C * k) @ kernel.reshape(kernel.shape[0], -1).T
    return out.reshape(N, L, -1).swapaxes(1, 2)


def max_pool1d(x: Tensor, kernel_size: int, stride: int, padding: int = 0):
    # reduces the LAST axis of the (N, C, k, n_out) window tensor exactly like the reference (functional.py:143-166),
    # i.e. the result is (N, C, k): the maximum over window positions for each in-window offset.
    return _im2col(_pad_nd(x, padding, 1), kernel_size, stride, 1).max(-1)


def avg_pool1d(x: Tensor, kernel_size: int, stride: int, padding: int = 0):
    return _im2col(_pad_nd(x, padding, 1), kernel_size, stride, 1).mean(-1)


def conv2d(x: Tensor, kernel: Tensor, padding: int = 0, stride: int = 1):
    """im2col convolution, square kernel / int stride / int padding (reference functional.py:254-281)."""
    if _fused.usable(x, kernel, op='conv2d'):
        return _fused.conv2d(x, kernel, padding, stride)
    N = x.shape[0]
    O, _, k, _ = kernel.shape
    col = _im2col(_pad_nd(x, padding, 2), k, stride, 2)
    oh, ow = col.shape[-2:]
    col = col.transpose(0, 4, 5, 1, 2, 3).reshape(N * oh * ow, -1)
    out = col @ kernel.reshape(O, -1).T
    return out.reshape(N, oh, ow, -1).transpose(0, 3, 1, 2)


def _pool2d(x, kernel_size, stride, padding, mode):
    if _fused.usable(x, op='pool2d'):
        return _fused.pool2d(x, kernel_size, stride, padding, mode)
    N, C = x.shape[:2]
    col = _im2col(_pad_nd(x, padding, 2), kernel_size, stride, 2)
    oh, ow = col.shape[-2:]
    col = col.transpose(0, 4, 5, 1, 2, 3).reshape(-1, kernel_size * kernel_size)
    out = col.max(1) if mode == "max" else col.mean(1)
    return out.reshape(N, oh, ow, C).transpose(0, 3, 1, 2)


def max_pool2d(x: Tensor, kernel_size: int, stride: int, padding=0):
    return _pool2d(x, kernel_size, stride, padding, "max")


def avg_pool2d(x: Tensor, kernel_size: int, stride: int, padding=0):
    return _pool2d(x, kernel_size, stride, padding, "avg")


# ---------------------------------------------------------------------- losses -----------------------
def _reduce(t, reduction):
    if reduction == 'mean':
        return tensor.mean(t)
    elif reduction == 'sum':
        return tensor.sum(t)
    raise ValueError("reduction must be mean or sum.")


def mse_loss(y_pred, y_true, reduction='mean'):
    return _reduce(function.square(y_pred - y_true), reduction)


def nll_loss(y_pred, y_true, reduction='mean'):
    return _reduce(-y_pred * y_true, reduction)


def cross_entropy_loss(y_pred, y_true, reduction='mean'):
    """Log-sum-exp over axis 1; integer targets pick one entry per row, one-hot targets weight all N*C entries (so
    'mean' divides by N*C there) — reference functional.py:364-381."""
    if reduction not in ('mean', 'sum'):
        raise ValueError("reduction must be mean or sum.")
    if y_true.ndim == 1 and y_pred.ndim == 2 and _fused.usable(y_pred, op='cross_entropy'):
        return _fused.cross_entropy(y_pred, y_true, reduction)
    shifted = y_pred - y_pred.max().item()
    neg_log_sm = tensor.log(tensor.sum(tensor.exp(shifted), 1, keepdims=True)) - shifted
    if y_true.ndim == 1:
        nll = neg_log_sm[range(len(neg_log_sm)), y_true]
    else:
        nll = neg_log_sm * y_true
    return _reduce(nll, reduction)
