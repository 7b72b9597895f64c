"""Flat parameter / gradient / moment storage for the fused multi-tensor Adam kernel and the data-parallel bucket."""
import numpy as np

from ..backend import lib as L
from ..backend.array import ndarray


def _align(n, a=64):
    return (n + a - 1) // a * a


class FlatAdamState:
    """Moves every parameter's storage into one flat fp32 buffer (parameters keep their shapes as views — user code that
    writes ``p.data[...] = w`` keeps working) and pins every ``.grad`` to a view of a second flat buffer."""

    def __init__(self, params):
        self.params = params
        dev = params[0].device
        offs, total = [], 0
        for p in params:
            offs.append(total)
            total += _align(p.size)
        self.total = total
        with dev:
            self.device = dev
            self.flat_p = ndarray.empty((total, ), np.float32)
            self.flat_g = ndarray.empty((total, ), np.float32)
            self.flat_m = ndarray.empty((total, ), np.float32)
            self.flat_v = ndarray.empty((total, ), np.float32)
            for buf in (self.flat_p, self.flat_g, self.flat_m, self.flat_v):
                buf.fill(0.0)
            self.m_views, self.v_views = [], []
            for p, o in zip(params, offs):
                def view(buf):
                    return buf._view((p.size, ), (1, ), o).reshape(p.shape)
                pv = view(self.flat_p)
                pv[...] = p.data
                p.data = pv
                gv = view(self.flat_g)
                if p._grad is not None and not p._grad_stale:
                    gv[...] = p._grad
                    p._grad = gv
                    p._grad_stale = False
                else:
                    p._grad = gv
                    p._grad_stale = True
                p._pinned_grad = True
                p._grad_dtype = np.dtype(np.float32)
                self.m_views.append(view(self.flat_m))
                self.v_views.append(view(self.flat_v))

    def _settle_grads(self):
        """A parameter that received no gradient since zero_grad() holds stale values: its grad is zero."""
        for p in self.params:
            if p._grad_stale:
                p._grad.fill(0.0)
                p._grad_stale = False

    def step(self, lr, b1, b2, eps, wd, t, grad_scale):
        with self.device:
            self._settle_grads()
            self.flat_p.buf.version += 1
            L.call("pdn_adam_step", self.flat_p.ptr, self.flat_g.ptr, self.flat_m.ptr, self.flat_v.ptr, self.total, lr, b1, b2, eps,
                   wd, t, grad_scale)
