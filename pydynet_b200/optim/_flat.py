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
        self.total, self.offs = total, offs
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
                pv = self._view(self.flat_p, p, o)
                pv[...] = p.data
                p.data = pv
                gv = self._view(self.flat_g, p, o)
                if p._grad is not None and not p._grad_stale:
                    gv[...] = p._grad
                    p._grad = gv
                    p._grad_stale = False
                else:
                    p._grad = gv
                    p._grad_stale = True
                p._pinned_grad = True
                p._grad_dtype = np.dtype(np.float32)
                self.m_views.append(self._view(self.flat_m, p, o))
                self.v_views.append(self._view(self.flat_v, p, o))

    @staticmethod
    def _view(buf, p, o):
        return buf._view((p.size, ), (1, ), o).reshape(p.shape)

    def repin(self):
        """User code may re-bind a parameter's storage after the optimizer was built (``p.data = w``, ``p.grad = g`` on an un-pinned
        tensor, ``Module.to`` to the same device): the fused kernel and the data-parallel bucket only see the flat buffers, so a
        stray array is copied back into its segment and the parameter re-pinned. A parameter that left the device is an error."""
        for p, o in zip(self.params, self.offs):
            d = p.data
            if getattr(d, "buf", None) is not self.flat_p.buf:
                if p.device != self.device:
                    raise RuntimeError("a parameter was moved to another device after its optimizer was created; build a new optimizer")
                pv = self._view(self.flat_p, p, o)
                pv[...] = d
                p.data = pv
                self.flat_p.buf.version += 1
            g = p._grad
            if g is None or getattr(g, "buf", None) is not self.flat_g.buf:
                gv = self._view(self.flat_g, p, o)
                if g is not None and not p._grad_stale:
                    gv[...] = g
                    p._grad_stale = False
                else:
                    p._grad_stale = True
                p._grad = gv
                p._pinned_grad = True
                p._grad_dtype = np.dtype(np.float32)

    def _settle_grads(self):
        """A parameter that received no gradient since zero_grad() holds stale values: its grad is zero."""
        for p in self.params:
            if p._grad_stale:
                p._grad.fill(0.0)
                p._grad_stale = False

    def step(self, lr, b1, b2, eps, wd, t, grad_scale):
        from .. import cuda
        with self.device:
            self.repin()
            self._settle_grads()
            self.flat_p.buf.version += 1
            if cuda.is_capturing():
                # a step being RECORDED (cuda.graphed_step): the step counter and the learning rate live in device memory and the
                # bias correction is computed there, so every replay of the recording is the NEXT optimizer step
                st = self.device_state(t, lr)
                L.call("pdn_adam_step_dev", self.flat_p.ptr, self.flat_g.ptr, self.flat_m.ptr, self.flat_v.ptr, self.total, b1, b2, eps, wd,
                       grad_scale, st["t"].ptr, st["lr"].ptr, st["step"].ptr)
                return
            L.call("pdn_adam_step", self.flat_p.ptr, self.flat_g.ptr, self.flat_m.ptr, self.flat_v.ptr, self.total, lr, b1, b2, eps,
                   wd, t, grad_scale)

    _dev_state = None

    def device_state(self, t=None, lr=None):
        """{t: int32[1], lr: float32[1], step: float32[1]} on the device (created outside a recording by cuda.graphed_step)."""
        if self._dev_state is None:
            if t is None:
                return None
            from .. import cuda
            assert not cuda.is_capturing(), "the device-side optimizer state must exist before the step is recorded"
            with self.device:
                self._dev_state = {"t": ndarray.from_host(np.array([t], np.int32)), "lr": ndarray.from_host(np.array([lr], np.float32)),
                                   "step": ndarray.from_host(np.zeros(1, np.float32)), "lr_host": float(lr)}
        return self._dev_state

    def sync_device_state(self, t, lr):
        """Host -> device before a recording or when the host values were changed behind its back (lr schedulers)."""
        st = self.device_state(t, lr)
        with self.device:
            st["t"][...] = np.array([t], np.int32)
            st["lr"][...] = np.array([lr], np.float32)
            st["lr_host"] = float(lr)
