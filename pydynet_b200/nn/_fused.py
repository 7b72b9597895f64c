"""Fused autograd nodes for fp32 cuda tensors: one hand-written kernel (or a short fixed sequence) per functional op,
each a single tape entry.  Definitions of the results are the operator chains in nn/functional.py and nn/modules/*;
this module only changes how many launches and HBM passes they cost.  There is no host path in here: every function
raises if the backend library is missing."""
import ctypes as C
import os

import numpy as np
from ..backend.array import WRITE_EPOCH

from ..autograd import is_grad_enable
from ..core.tensor import Tensor, _result, _sum_to

_ENABLED = os.environ.get("PDN_FUSED", "1") != "0"
F32 = np.dtype(np.float32)


_HAVE = set()


def fused_op(fn):
    """Registers a fused implementation under its function name."""
    _HAVE.add(fn.__name__)
    return fn


def usable(*tensors, op=None) -> bool:
    """True when fused op ``op`` exists, fusion is enabled and every given tensor (None entries skipped) is an fp32
    cuda tensor."""
    if not _ENABLED or op not in _HAVE:
        return False
    for t in tensors:
        if t is None:
            continue
        if not isinstance(t, Tensor) or not t.device.is_cuda or t.data.dtype != F32:
            return False
    return True


def _bk():
    from .. import backend
    return backend


def _call(name, *args):
    from ..backend import lib
    lib.call(name, *args)


def _c(a):
    """C-contiguous fp32 device array."""
    return a if a.is_contiguous else a.copy()


def _empty(shape):
    from ..backend.array import ndarray
    return ndarray.empty(shape, F32)


def _needs(*ts):
    return is_grad_enable() and any(t is not None and t.requires_grad for t in ts)


# ---------------------------------------------------------------------------------- linear -------------------
class _PackedWeight:
    """bf16 hi/lo operand planes of a constant weight matrix, owned by libpdn_b200 (pdn_gemm_prepack)."""
    __slots__ = ("handle", "key")

    def __init__(self, w):
        h = C.c_void_p()
        K, N = w.shape
        _call("pdn_gemm_prepack", w.ptr, K, N, w.estrides[0], w.estrides[1], C.byref(h))
        self.handle = h
        self.key = (w.ptr, w.buf.version, w.shape, w.estrides)

    def __del__(self):
        try:
            from ..backend import lib
            if lib._lib is not None and self.handle:
                lib._lib.pdn_gemm_prepack_free(self.handle)
        except Exception:
            pass


def _packed(weight):
    w = weight.data
    pw = getattr(weight, "_pdn_packed", None)
    if pw is None or pw.key != (w.ptr, w.buf.version, w.shape, w.estrides):
        pw = _PackedWeight(w)
        weight._pdn_packed = pw
    return pw


@fused_op
def linear(x, weight, bias):
    """x @ W + b as one GEMM with the bias added in the epilogue (F.linear, reference functional.py:7-11); backward
    contracts dW over all leading dims at once and reduces db with one column-sum. Under no_grad (inference) the weight's
    tcgen05 operand planes are packed once and cached until the weight buffer is written again."""
    bk = _bk()
    with x.device:
        xd = x.data
        lead = xd.shape[:-1]
        x2 = bk.ext._flat2d(xd) if xd.ndim != 2 else xd
        M, N = x2.shape[0], weight.shape[1]
        bd = _c(bias.data) if bias is not None else None
        if not is_grad_enable() and M >= 32 and weight.data.ndim == 2:
            out = _empty((M, N))
            _call("pdn_gemm_prepacked", x2.ptr, _packed(weight).handle, out.ptr, M, x2.estrides[0], x2.estrides[1], N,
                  bd.ptr if bd is not None else None, 0)
        else:
            out = bk.gemm_into(None, x2, weight.data, bias=bd)
        data = out.reshape(lead + (N, )) if xd.ndim != 2 else out

    def backward(g):
        g2 = bk.ext._flat2d(g) if g.ndim != 2 else g
        gx = gw = gb = None
        if x.requires_grad:
            gx = bk.gemm_into(None, g2, weight.data.swapaxes(0, 1))
            gx = gx.reshape(x.shape) if xd.ndim != 2 else gx
        if weight.requires_grad:
            gw = bk.gemm_into(None, x2.swapaxes(0, 1), g2)
        if bias is not None and bias.requires_grad:
            gb = g2.sum(axis=0)
        return gx, gw, gb

    ins = (x, weight) + ((bias, ) if bias is not None else ())
    return _result(data, x.device, ins, (lambda g: backward(g)[:len(ins)]), "linear")


# ---------------------------------------------------------------------------------- rows ---------------------
@fused_op
def softmax(x, log=False):
    """softmax / log_softmax over the last axis: one kernel forward, one backward (reference functional.py:43-58)."""
    with x.device:
        xd = _c(x.data)
        n = xd.shape[-1]
        rows = xd.size // n if n else 0
        y = _empty(xd.shape)
        _call("pdn_softmax_fwd", 0, xd.ptr, y.ptr, rows, n, int(log))

    def backward(g):
        g = _c(g)
        dx = _empty(xd.shape)
        _call("pdn_softmax_bwd", 0, y.ptr, g.ptr, dx.ptr, rows, n, int(log))
        return (dx, )

    return _result(y, x.device, (x, ), backward, "log_softmax" if log else "softmax")


@fused_op
def rmsnorm(x, weight, eps):
    with x.device:
        xd, w = _c(x.data), _c(weight.data)
        n = xd.shape[-1]
        rows = xd.size // n
        y, rstd = _empty(xd.shape), _empty((rows, ))
        _call("pdn_rmsnorm_fwd", xd.ptr, w.ptr, y.ptr, rstd.ptr, rows, n, eps)

    def backward(g):
        g = _c(g)
        dx = _empty(xd.shape) if x.requires_grad else None
        dw = _empty((n, )) if weight.requires_grad else None
        _call("pdn_rmsnorm_bwd", xd.ptr, w.ptr, rstd.ptr, g.ptr, dx.ptr if dx is not None else None, dw.ptr if dw is not None else None,
              rows, n)
        return dx, dw

    return _result(y, x.device, (x, weight), backward, "rmsnorm")


@fused_op
def swiglu(gate, up):
    """silu(gate) * up (reference llm/llama/model.py:56-58) in one pass."""
    with gate.device:
        a, b = _c(gate.data), _c(up.data)
        out = _empty(a.shape)
        _call("pdn_swiglu", a.ptr, b.ptr, out.ptr, a.size)

    def backward(g):
        g = _c(g)
        da, db = _empty(a.shape), _empty(a.shape)
        _call("pdn_swiglu_bwd", a.ptr, b.ptr, g.ptr, da.ptr, db.ptr, a.size)
        return da, db

    return _result(out, gate.device, (gate, up), backward, "swiglu")


@fused_op
def silu(x):
    from ..backend import lib as L
    from ..backend.array import _unary, ternary
    with x.device:
        out = _unary(L.SILU, x.data)
    return _result(out, x.device, (x, ), lambda g: (ternary(L.T_SILU_GRAD, x.data, g, g), ), "silu")


@fused_op
def cross_entropy(logits, target, reduction):
    """log-sum-exp + target pick + mean/sum in one kernel (reference functional.py:364-381, integer targets)."""
    bk = _bk()
    with logits.device:
        x = _c(logits.data)
        N, Cn = x.shape
        t = target.data if target.data.dtype == np.int64 else target.data.astype(np.int64)
        t = _c(t)
        loss, lse = _empty(()), _empty((N, ))
        mean = int(reduction == "mean")
        _call("pdn_ce_loss_fwd", x.ptr, t.ptr, loss.ptr, lse.ptr, N, Cn, mean)

    def backward(g):
        g = _c(g if g.dtype == F32 else g.astype(F32))
        dx = _empty(x.shape)
        _call("pdn_ce_loss_bwd", x.ptr, t.ptr, lse.ptr, g.ptr, dx.ptr, N, Cn, mean)
        return (dx, )

    return _result(loss, logits.device, (logits, ), backward, "cross_entropy")


# ---------------------------------------------------------------------------------- attention ----------------
_ATT_FFMA_LIMIT = 1 << 29  # B*H*Lq*Lk*D multiply-adds served by the one-warp-per-query kernel; larger -> GEMM path


_ATT_TC_MIN = 1 << 22  # below this many multiply-adds the warp-per-query kernel's single launch wins


def _attention_impl(B, H, Lq, Lk, D) -> str:
    """'tc' (tcgen05 flash kernels, D <= 64), 'ffma' (warp-per-query kernel, D <= 128) or '' (neither: composite path)."""
    forced = os.environ.get("PDN_ATTN", "")
    work = B * H * Lq * Lk * D
    if forced == "tc" and D <= 64:
        return "tc"
    if forced == "ffma" and D <= 128:
        return "ffma"
    if D <= 64 and work >= _ATT_TC_MIN and B * H <= 65535:
        return "tc"
    if D <= 128 and work <= _ATT_FFMA_LIMIT:
        return "ffma"
    return ""


def attention_fits(xq, xk) -> bool:
    B, Lq, H, D = xq.shape
    return _attention_impl(B, H, Lq, xk.shape[1], D) != ""


def _i64(vals):
    return (C.c_int64 * len(vals))(*[int(v) for v in vals])


def _bhl_strides(a):
    """(batch, head, row) element strides of a [B, L, H, D] array whose D axis is unit-stride."""
    return _i64((a.estrides[0], a.estrides[2], a.estrides[1]))


def _mask_args(mask, B, H, Lq, Lk):
    if mask is None:
        return None, None, None
    m = mask.data if isinstance(mask, Tensor) else mask
    if m.dtype != F32:
        m = m.astype(F32)
    mv = m.broadcast_to((B, H, Lq, Lk))
    if mv.estrides[1] != 0 and H > 1 or (Lk > 1 and mv.estrides[3] != 1):
        m = mv.copy()[:, 0]  # per-head masks are not supported by the kernel; none of the reference models use them
        mv = m.broadcast_to((B, H, Lq, Lk))
    return mv, mv.ptr, _i64((mv.estrides[0] if B > 1 else 0, mv.estrides[2] if Lq > 1 else 0))


def _unit_last(a):
    return a if (a.shape[-1] == 1 or a.estrides[-1] == 1) else a.copy()


@fused_op
def attention(xq, xk, xv, mask, scale):
    """softmax(q kᵀ·scale + mask) v for [B, L, H, D] head-split views; returns [B, Lq, H*D]
    (llm/llama/model.py:112-121, examples/pydynet/transformer.py:93-104). One fused operator forward, one backward;
    training-sized problems run on the tensor cores (csrc/attention_tc.cu), small ones on the warp-per-query kernel."""
    with xq.device:
        q, k, v = _unit_last(xq.data), _unit_last(xk.data), _unit_last(xv.data)
        B, Lq, H, D = q.shape
        Lk = k.shape[1]
        impl = _attention_impl(B, H, Lq, Lk, D)
        assert impl, "attention: shape not supported by the fused kernels (check attention_fits first)"
        out, lse = _empty((B, Lq, H, D)), _empty((B, H, Lq))
        keep, mptr, mstr = _mask_args(mask, B, H, Lq, Lk)
        vers = _i64((q.buf.version, k.buf.version, v.buf.version)) if is_grad_enable() else None  # backward re-uses the packs
        if impl == "tc":
            _call("pdn_attention_tc_fwd", q.ptr, k.ptr, v.ptr, mptr, out.ptr, lse.ptr, B, H, Lq, Lk, D, _bhl_strides(q), _bhl_strides(k),
                  _bhl_strides(v), mstr, scale, vers)
        else:
            _call("pdn_attention_fwd", q.ptr, k.ptr, v.ptr, mptr, out.ptr, lse.ptr, B, H, Lq, Lk, D, _bhl_strides(q), _bhl_strides(k),
                  _bhl_strides(v), mstr, scale, None, 0)

    def backward(g):
        g = _c(g.reshape(B, Lq, H, D))
        dq = _empty((B, Lq, H, D)) if xq.requires_grad else None
        dk = _empty((B, Lk, H, D)) if xk.requires_grad else None
        dv = _empty((B, Lk, H, D)) if xv.requires_grad else None
        extra = (_i64((q.buf.version, k.buf.version, v.buf.version)), ) if impl == "tc" else ()
        _call("pdn_attention_tc_bwd" if impl == "tc" else "pdn_attention_bwd", q.ptr, k.ptr, v.ptr, mptr, out.ptr, lse.ptr, g.ptr,
              dq.ptr if dq is not None else None, dk.ptr if dk is not None else None, dv.ptr if dv is not None else None, B, H, Lq, Lk, D,
              _bhl_strides(q), _bhl_strides(k), _bhl_strides(v), mstr, scale, *extra)
        _ = keep
        return dq, dk, dv

    return _result(out.reshape(B, Lq, H * D), xq.device, (xq, xk, xv), backward, "attention")


class Planes:
    """A GEMM A-operand already in tensor-core format: bf16 hi/lo planes [2][M][Kp] written by a producer kernel."""
    __slots__ = ("buf", "M", "K", "Kp", "lead", "device")

    def __init__(self, lead, K, device):
        from ..backend.array import ndarray
        self.lead, self.K, self.device = tuple(lead), K, device
        self.M = int(np.prod(lead, dtype=np.int64)) if lead else 1
        self.Kp = (K + 7) // 8 * 8
        self.buf = ndarray.empty((4 * self.M * self.Kp, ), np.uint8)

    @property
    def ptr(self):
        return self.buf.ptr


def rmsnorm_planes(x, weight, eps):
    """RMSNorm whose result feeds only GEMMs: emitted directly as operand planes (inference)."""
    with x.device:
        xd, w = _c(x.data), _c(weight.data)
        n = xd.shape[-1]
        pl = Planes(xd.shape[:-1], n, x.device)
        _call("pdn_rmsnorm_planes", xd.ptr, w.ptr, pl.ptr, pl.M, n, pl.Kp, eps)
    return pl


def linear_residual_(a, weight, res):
    """res += a @ W, accumulated in the GEMM epilogue INTO res's buffer (inference only: the residual stream is not
    needed in its old state). Returns res."""
    bk = _bk()
    with a.device:
        rd = res.data
        assert rd.is_contiguous and rd.dtype == F32
        N = weight.shape[1]
        rd.buf.version += 1
        if isinstance(a, Planes):
            _call("pdn_gemm_prepacked_planes", a.ptr, a.M, a.Kp, _packed(weight).handle, rd.ptr, N, None, 1)
            return res
        ad = a.data
        a2 = bk.ext._flat2d(ad) if ad.ndim != 2 else ad
        _call("pdn_gemm_prepacked", a2.ptr, _packed(weight).handle, rd.ptr, a2.shape[0], a2.estrides[0], a2.estrides[1], N, None, 1)
    return res


class DevicePos:
    """Sequence position of a decode step held in DEVICE memory (an int64 [1] tensor) so that a CUDA-graph recording of
    the step stays valid while the position advances."""

    def __init__(self, tensor):
        self.tensor = tensor


# ---------------------------------------------------------------------------------- conv / pool --------------
@fused_op
def conv2d(x, kernel, padding, stride, bias=None):
    """F.conv2d (+ Conv2d's (1,O,1,1) bias) as gathered tcgen05 GEMMs (csrc/conv.cu); NCHW contiguous result
    (reference functional.py:254-281, conv.py:99-103)."""
    with x.device:
        xd, wd = _c(x.data), _c(kernel.data)
        N, Cin, H, W = xd.shape
        O, _, k, _ = wd.shape
        oh, ow = (H + 2 * padding - k) // stride + 1, (W + 2 * padding - k) // stride + 1
        y = _empty((N, O, oh, ow))
        bd = _c(bias.data).reshape(-1) if bias is not None else None
        xver = xd.buf.version if is_grad_enable() else -1  # backward-weight re-uses the forward's operand planes of x
        _call("pdn_conv2d_fwd", xd.ptr, wd.ptr, bd.ptr if bd is not None else None, y.ptr, N, Cin, H, W, O, k, stride, padding, xver)

    def backward(g):
        g = _c(g)
        dx = dw = db = None
        if x.requires_grad:
            dx = _empty(xd.shape)
            _call("pdn_conv2d_bwd_data", g.ptr, wd.ptr, dx.ptr, N, Cin, H, W, O, k, stride, padding, g.buf.version)
        need_b = bias is not None and bias.requires_grad
        if kernel.requires_grad or need_b:
            dw = _empty(wd.shape) if kernel.requires_grad else None
            db = _empty(bias.shape) if need_b else None
            _call("pdn_conv2d_bwd_weight", xd.ptr, g.ptr, dw.ptr if dw is not None else None, db.ptr if db is not None else None, N, Cin, H,
                  W, O, k, stride, padding, xd.buf.version, g.buf.version)
        return (dx, dw, db) if bias is not None else (dx, dw)

    ins = (x, kernel) + ((bias, ) if bias is not None else ())
    return _result(y, x.device, ins, backward, "conv2d")


@fused_op
def pool2d(x, k, stride, padding, mode):
    """max / avg pooling as a direct window kernel; zero padding participates, tied maxima all receive the gradient
    (reference functional.py:284-339, tensor.py:741-747)."""
    m = 0 if mode == "max" else 1
    with x.device:
        xd = _c(x.data)
        N, Cn, H, W = xd.shape
        oh, ow = (H + 2 * padding - k) // stride + 1, (W + 2 * padding - k) // stride + 1
        y = _empty((N, Cn, oh, ow))
        _call("pdn_pool2d_fwd", xd.ptr, y.ptr, N, Cn, H, W, k, stride, padding, m)

    def backward(g):
        g = _c(g)
        dx = _empty(xd.shape)
        _call("pdn_pool2d_bwd", xd.ptr, y.ptr, g.ptr, dx.ptr, N, Cn, H, W, k, stride, padding, m)
        return (dx, )

    return _result(y, x.device, (x, ), backward, "pool2d")


# ---------------------------------------------------------------------------------- feature-statistic norm ----
@fused_op
def feature_norm(mod, x, axes, keep):
    """Training-mode forward of BatchNorm1d/2d and the reference's batch-statistic "LayerNorm" (norm.py:58-73, 132-147,
    203-218) incl. the in-place running-stat update; 3 kernels forward, 3 backward instead of ~12 eager nodes."""
    scale, shift = mod.scale, mod.shift
    red = (axes, ) if isinstance(axes, int) else tuple(axes)
    with x.device:
        xd = _c(x.data)
        nd = xd.ndim
        if red == tuple(range(len(red))):  # leading axes reduced: [outer, C]
            outer = int(np.prod(xd.shape[:len(red)], dtype=np.int64))
            Cn, inner = xd.size // max(outer, 1), 1
        elif red == (0, ) + tuple(range(2, nd)):  # channel axis 1: [N, C, inner]
            outer, Cn, inner = xd.shape[0], xd.shape[1], int(np.prod(xd.shape[2:], dtype=np.int64))
        else:
            raise NotImplementedError(f"feature_norm over axes {red}")
        stat_shape = scale.shape
        mean, var = _empty((Cn, )), _empty((Cn, ))
        from .. import distributed as dist
        sync = dist.sync_stats_enabled()
        W = dist.get_world_size() if sync else 1
        inv_mg = 1.0 / (outer * inner * W)  # 1 / (elements per feature in the GLOBAL batch)
        if sync:  # partial sums pre-scaled by 1/m_global, summed across ranks on the compute stream
            _call("pdn_bnorm_partial", xd.ptr, None, mean.ptr, outer, Cn, inner, 0, inv_mg)
            dist.all_reduce_sum_(mean)
            _call("pdn_bnorm_partial", xd.ptr, mean.ptr, var.ptr, outer, Cn, inner, 1, inv_mg)
            dist.all_reduce_sum_(var)
        else:
            _call("pdn_bnorm_stats", xd.ptr, mean.ptr, var.ptr, outer, Cn, inner)
        y = _empty(xd.shape)
        sc, sh = _c(scale.data).reshape(-1), _c(shift.data).reshape(-1)
        _call("pdn_bnorm_apply", xd.ptr, mean.ptr, var.ptr, sc.ptr, sh.ptr, y.ptr, outer, Cn, inner, mod.eps)
        rm, rv = mod.running_mean.data, mod.running_var.data
        if rm.dtype == F32 and rv.dtype == F32 and rm.size == Cn and rv.size == Cn and rm.is_contiguous and rv.is_contiguous:
            _call("pdn_bnorm_running", rm.ptr, rv.ptr, mean.ptr, var.ptr, float(mod.momentum), Cn)  # one launch instead of six
            rm.buf.version += 1
            rv.buf.version += 1
            WRITE_EPOCH[0] += 1
        else:
            rm *= (1 - mod.momentum)
            rm += (mean * mod.momentum).reshape(rm.shape)
            rv *= (1 - mod.momentum)
            rv += (var * mod.momentum).reshape(rv.shape)

    def backward(g):
        g = _c(g)
        dx = _empty(xd.shape) if x.requires_grad else None
        dsc, dsh = _empty((Cn, )), _empty((Cn, ))
        if sync:
            mg, mgx = _empty((Cn, )), _empty((Cn, ))
            _call("pdn_bnorm_bwd_reduce", xd.ptr, mean.ptr, var.ptr, g.ptr, mg.ptr, mgx.ptr, outer, Cn, inner, mod.eps, inv_mg)
            dsh, dsc = mg * (1.0 / inv_mg), mgx * (1.0 / inv_mg)  # this rank's share of dshift / dscale
            dist.all_reduce_sum_(mg)
            dist.all_reduce_sum_(mgx)
            if dx is not None:
                _call("pdn_bnorm_bwd_dx", xd.ptr, mean.ptr, var.ptr, sc.ptr, g.ptr, mg.ptr, mgx.ptr, dx.ptr, outer, Cn, inner, mod.eps)
        else:
            _call("pdn_bnorm_bwd", xd.ptr, mean.ptr, var.ptr, sc.ptr, g.ptr, dx.ptr if dx is not None else None, dsc.ptr, dsh.ptr, outer, Cn,
                  inner, mod.eps)
        return dx, dsc.reshape(stat_shape), dsh.reshape(stat_shape)

    return _result(y, x.device, (x, scale, shift), backward, "feature_norm")


# ---------------------------------------------------------------------------------- recurrent sequences ------
def _seq_inputs(x):
    """[T, B, I] device array (unbatched [T, I] gets B = 1) flattened to contiguous [T*B, I]."""
    xd = x.data
    if xd.ndim == 2:
        xd = xd.reshape(xd.shape[0], 1, xd.shape[1])
    T, B, I = xd.shape
    return _c(xd).reshape(T * B, I), T, B, I


def sequence(cell, x, state):
    """Whole-sequence forward of one recurrent layer/direction as ONE tape entry (csrc/rnn.cu). Returns
    ([state sequences], [last states]) like the per-step loop in nn/modules/rnn.py, or None when not applicable."""
    from .modules.rnn import GRUCell, LSTMCell, RNNCell
    if not _ENABLED or type(cell) not in (GRUCell, LSTMCell, RNNCell) or x.ndim not in (2, 3):
        return None
    tensors = [x] + list(state) + list(cell._parameters.values())
    if not all(t.device.is_cuda and t.data.dtype == F32 for t in tensors):
        return None
    if type(cell) is RNNCell:
        return _rnn_sequence(cell, x, state[0]) if cell.nonlinearity in ("tanh", "relu") else None
    return _gru_sequence(cell, x, state[0]) if type(cell) is GRUCell else _lstm_sequence(cell, x, state[0], state[1])


def _rnn_sequence(cell, x, h0):
    """h_t = act(x_t Wx + b + h_{t-1} Wh) over a whole sequence (reference rnn.py:38-49 per step): hoisted input GEMM, the
    recurrence and its BPTT as one C call each (pdn_rnn_seq_fwd / _bwd)."""
    bk = _bk()
    H = cell.hidden_size
    Wx, Wh = cell.Wx, cell.Wh
    b = cell.bias if cell.has_bias else None
    relu = 1 if cell.nonlinearity == "relu" else 0
    unb = x.ndim == 2
    with x.device:
        x2, T, B, I = _seq_inputs(x)
        h0d = _c(h0.data).reshape(B, H)
        xp = bk.gemm_into(None, x2, Wx.data, bias=_c(b.data) if b is not None else None)
        hs = _empty((T, B, H))
        wh = _c(Wh.data)
        _call("pdn_rnn_seq_fwd", xp.ptr, h0d.ptr, wh.ptr, hs.ptr, T, B, H, relu)
        del xp

    def backward(g):
        g = _c(g.reshape(T, B, H))
        dxp, dh0, dWh = _empty((T * B, H)), _empty((B, H)), _empty((H, H))
        _call("pdn_rnn_seq_bwd", g.ptr, h0d.ptr, hs.ptr, wh.ptr, dxp.ptr, dh0.ptr, dWh.ptr, T, B, H, relu)
        dx = bk.gemm_into(None, dxp, Wx.data.swapaxes(0, 1)).reshape(x.shape) if x.requires_grad else None
        outs = [dx, dh0.reshape(h0.shape), bk.gemm_into(None, x2.swapaxes(0, 1), dxp), dWh]
        if b is not None:
            outs.append(dxp.sum(axis=0))
        return tuple(outs)

    ins = (x, h0, Wx, Wh) + ((b, ) if b is not None else ())
    seq = _result(hs.reshape(T, H) if unb else hs, x.device, ins, backward, "rnn_sequence")
    return [seq], [seq[T - 1:T]]


def _gru_sequence(cell, x, h0):
    bk = _bk()
    H = cell.hidden_size
    Wx1, Wh1, Wx2, Wh2 = cell.Wx1, cell.Wh1, cell.Wx2, cell.Wh2
    b1, b2 = (cell.bias1, cell.bias2) if cell.has_bias else (None, None)
    unb = x.ndim == 2
    with x.device:
        x2, T, B, I = _seq_inputs(x)
        h0d = _c(h0.data).reshape(B, H)
        xp1 = bk.gemm_into(None, x2, Wx1.data, bias=_c(b1.data) if b1 is not None else None)  # hoisted over all T
        xp2 = bk.gemm_into(None, x2, Wx2.data, bias=_c(b2.data) if b2 is not None else None)
        hs, zr, nn_ = _empty((T, B, H)), _empty((T, B, 2 * H)), _empty((T, B, H))
        w1, w2 = _c(Wh1.data), _c(Wh2.data)
        _call("pdn_gru_seq_fwd", xp1.ptr, xp2.ptr, h0d.ptr, w1.ptr, w2.ptr, hs.ptr, zr.ptr, nn_.ptr, T, B, H)
        del xp1, xp2

    def backward(g):
        g = _c(g.reshape(T, B, H))
        dxp1, dxp2 = _empty((T * B, 2 * H)), _empty((T * B, H))
        dh0, dW1, dW2 = _empty((B, H)), _empty((H, 2 * H)), _empty((H, H))
        _call("pdn_gru_seq_bwd", g.ptr, h0d.ptr, hs.ptr, zr.ptr, nn_.ptr, w1.ptr, w2.ptr, dxp1.ptr, dxp2.ptr, dh0.ptr, dW1.ptr, dW2.ptr, T, B, H)
        dx = None
        if x.requires_grad:
            dx = bk.gemm_into(None, dxp1, Wx1.data.swapaxes(0, 1))
            bk.gemm_into(dx, dxp2, Wx2.data.swapaxes(0, 1), accumulate=True)
            dx = dx.reshape(x.shape)
        xt = x2.swapaxes(0, 1)
        outs = [dx, dh0.reshape(h0.shape), bk.gemm_into(None, xt, dxp1), dW1, bk.gemm_into(None, xt, dxp2), dW2]
        if b1 is not None:
            outs += [dxp1.sum(axis=0), dxp2.sum(axis=0)]
        return tuple(outs)

    ins = (x, h0, Wx1, Wh1, Wx2, Wh2) + ((b1, b2) if b1 is not None else ())
    seq = _result(hs.reshape(T, H) if unb else hs, x.device, ins, backward, "gru_sequence")
    return [seq], [seq[T - 1:T]]


def _lstm_sequence(cell, x, h0, c0):
    bk = _bk()
    H = cell.hidden_size
    Wx, Wh = cell.Wx, cell.Wh
    b = cell.bias if cell.has_bias else None
    unb = x.ndim == 2
    with x.device:
        x2, T, B, I = _seq_inputs(x)
        h0d, c0d = _c(h0.data).reshape(B, H), _c(c0.data).reshape(B, H)
        xp = bk.gemm_into(None, x2, Wx.data, bias=_c(b.data) if b is not None else None)
        hc, gates = _empty((2, T, B, H)), _empty((T, B, 4 * H))
        wh = _c(Wh.data)
        cs_ptr = hc.ptr + T * B * H * 4
        _call("pdn_lstm_seq_fwd", xp.ptr, h0d.ptr, c0d.ptr, wh.ptr, hc.ptr, cs_ptr, gates.ptr, T, B, H)
        del xp

    def backward(g):
        g = _c(g.reshape(2, T, B, H))
        g_cT = _c(g[1, T - 1])  # only the final cell state is exposed by the module API
        dxp, dh0, dc0, dWh = _empty((T * B, 4 * H)), _empty((B, H)), _empty((B, H)), _empty((H, 4 * H))
        _call("pdn_lstm_seq_bwd", g.ptr, g_cT.ptr, h0d.ptr, c0d.ptr, hc.ptr, cs_ptr, gates.ptr, wh.ptr, dxp.ptr, dh0.ptr, dc0.ptr, dWh.ptr, T, B, H)
        dx = bk.gemm_into(None, dxp, Wx.data.swapaxes(0, 1)).reshape(x.shape) if x.requires_grad else None
        outs = [dx, dh0.reshape(h0.shape), dc0.reshape(c0.shape), bk.gemm_into(None, x2.swapaxes(0, 1), dxp), dWh]
        if b is not None:
            outs.append(dxp.sum(axis=0))
        return tuple(outs)

    ins = (x, h0, c0, Wx, Wh) + ((b, ) if b is not None else ())
    both = _result(hc.reshape(2, T, H) if unb else hc, x.device, ins, backward, "lstm_sequence")
    hs, cs = both[0], both[1]
    return [hs, cs], [hs[T - 1:T], cs[T - 1:T]]
