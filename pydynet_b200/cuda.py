"""Device selection — the reference's only plugin seam (reference pydynet/cuda.py:16-99).

``Device.xp`` returns NumPy for ``cpu`` and :mod:`pydynet_b200.backend` (hand-written sm_100a kernels behind a
ctypes C ABI) for ``cuda`` — the slot CuPy occupies in the reference.  A cuda device that cannot be served
(no libpdn_b200.so, no GPU) raises ``RuntimeError`` like the reference does without CuPy (cuda.py:67-69);
nothing silently falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .backend import lib as _L


def is_available() -> bool:
    return _L.device_count() > 0


def __getattr__(name):
    # reference cuda.py:5-13 exposes a module-level flag (True when CuPy imported); here: a GPU that libpdn_b200.so can serve
    if name == "cuda_available":
        return is_available()
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def __dir__():
    return sorted(list(globals()) + ["cuda_available"])


def device_count() -> int:
    return _L.device_count()


_current = [None]  # device id this process last selected through this module (every switch goes through set_device)


def current_device() -> int:
    cur = _current[0]
    if cur is None:
        d = C.c_int(0)
        _L.call("pdn_get_device", C.byref(d))
        cur = _current[0] = int(d.value)
    return cur


def set_device(device: int) -> None:
    _L.call("pdn_set_device", int(device))
    _current[0] = int(device)


def synchronize() -> None:
    _L.call("pdn_sync")


class Device:
    __slots__ = ("device", "device_id", "_prev")

    def __init__(self, device=None) -> None:
        self._prev = None
        if isinstance(device, Device):
            self.device = device.device
            self.device_id = device.device_id
            return
        self.device_id = None
        if device is None:
            self.device = "cpu"
        elif isinstance(device, str):
            if device == "cpu":
                self.device = "cpu"
            elif device.startswith("cuda"):
                rest = device[4:]
                if rest == "":
                    rest = ":0"
                cuda_id = rest.split(":")[-1]
                if not rest.startswith(":") or not cuda_id.isdigit():
                    raise ValueError(f'Wrong cuda id "{cuda_id}"!')
                self.device = "cuda"
                self.device_id = int(cuda_id)
            else:
                raise ValueError(f'Unknown device "{device}"!')
        elif isinstance(device, (int, np.integer)):
            self.device = "cuda"
            self.device_id = int(device)
        else:
            raise ValueError(f"Unknown device {device!r}!")
        if self.device == "cuda":
            if not is_available():
                raise RuntimeError("Cuda device is not supported on this system.")
            if self.device_id >= device_count():
                raise RuntimeError(f"cuda:{self.device_id} requested but only {device_count()} device(s) present.")
            _ensure_init(self.device_id)

    def __repr__(self) -> str:
        if self.device == "cpu":
            return "Device(type='cpu')"
        return "Device(type='cuda', index={})".format(self.device_id)

    def __eq__(self, other) -> bool:
        if not isinstance(other, Device):
            other = Device(other)
        return self.device == other.device and self.device_id == other.device_id

    def __hash__(self):
        return hash((self.device, self.device_id))

    @property
    def is_cuda(self) -> bool:
        return self.device == "cuda"

    @property
    def xp(self):
        if self.device == "cpu":
            return np
        from . import backend
        return backend

    def __enter__(self):
        # Device objects are shared between tensors, and `with dev:` blocks nest (an optimizer step that touches a lazily created
        # gradient re-enters the same object): every entry pushes what it has to restore, every exit pops its own entry
        prev = None
        if self.device == "cuda":
            cur = current_device()
            if cur != self.device_id:
                prev = cur
                set_device(self.device_id)
        if self._prev is None:
            self._prev = [prev]
        else:
            self._prev.append(prev)
        return self

    def __exit__(self, *exc):
        prev = self._prev.pop() if self._prev else None
        if prev is not None:
            set_device(prev)


_inited = set()


def _ensure_init(dev_id: int):
    if dev_id not in _inited:
        cur = None
        try:
            cur = current_device()
        except Exception:
            pass
        _L.call("pdn_init", dev_id)
        _current[0] = None  # pdn_init selects dev_id: re-query
        if cur is not None and cur != dev_id and cur in _inited:
            set_device(cur)
        _inited.add(dev_id)


_capturing = False


def is_capturing() -> bool:
    """True while launches on the library stream are being recorded into a CUDA graph instead of executed."""
    return _capturing


class Graph:
    """CUDA-graph capture / replay of a launch sequence on the library's compute stream (pdn_graph_* in
    include/pdn_b200.h). Memory allocated while capturing stays reserved for the graph until it is destroyed, so the
    recorded kernels' buffers cannot be handed to unrelated work between replays.

        g = Graph(); g.begin(); ...launch work...; g.end(); g.launch(); g.launch()
    """

    def __init__(self):
        self._exec = None

    def begin(self):
        global _capturing
        _L.call("pdn_graph_begin")
        _capturing = True

    def end(self):
        global _capturing
        _capturing = False
        h = C.c_void_p()
        _L.call("pdn_graph_end", C.byref(h))
        self._exec = h

    def launch(self):
        _L.call("pdn_graph_launch", self._exec)

    def destroy(self):
        if self._exec is not None:
            _L.call("pdn_graph_destroy", self._exec)
            self._exec = None

    def __del__(self):
        try:
            if _L._lib is not None:
                self.destroy()
        except Exception:
            pass


NVTX = os.environ.get("PDN_NVTX") is not None


class nvtx_range:
    """``with nvtx_range("backward"): ...`` — a named NVTX range on the calling thread (no-op unless PDN_NVTX=1, so the hot path
    pays one attribute test). The engine brackets ``Tensor.backward``, ``Optimizer.step`` (Adam), the data-parallel gradient
    exchange and the inference plans with it."""
    __slots__ = ("name", )

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if NVTX:
            _L.call("pdn_nvtx_push", self.name.encode())
        return self

    def __exit__(self, *exc):
        if NVTX:
            _L.call("pdn_nvtx_pop")


class graphed_step:
    """Records ONE training (or inference) step — everything ``fn`` launches: forward, loss, ``zero_grad`` / ``backward``, the
    optimizer update — into a CUDA graph and replays it on every further call: one graph launch instead of the ~20 µs of Python
    per eager array expression (a LeNet step is 47 launches, the 512^3 matmul fwd+bwd of BASELINE config 1 is 9).

        step = pdn.cuda.graphed_step(lambda X, y: train_step(net, opt, X, y), optimizers=[opt])
        for X, y in loader:
            loss = step(X, y)          # Tensor holding the step's loss; read it (.item()) only when you need it

    Contract (what makes a recording valid): ``fn`` receives device Tensors of FIXED shapes / dtypes (the values of every call are
    copied into the recorded input buffers), performs no host read (``.item()``, ``.numpy()``, printing a tensor) and no host-side
    randomness (Dropout masks come from host ``np.random`` like the reference's), and returns a Tensor or a tuple of Tensors. The first
    ``warmup`` calls run eagerly (allocator pools and operand-plane caches reach their steady state), the next one is recorded, later
    ones are replays. Adam optimizers passed in ``optimizers`` keep their step counter and learning rate in device memory during
    replays (the host attributes ``t`` / ``lr`` stay in step; a learning rate changed by a scheduler is uploaded before the replay).
    Memory allocated while recording stays reserved for the graph. Not part of the reference's API (it has no graphs); the eager
    path is unchanged for code that does not opt in."""

    def __init__(self, fn, optimizers=(), warmup: int = 2):
        self.fn, self.optimizers, self.warmup = fn, list(optimizers), int(warmup)
        self.calls, self.graph, self.static_in, self.out = 0, None, None, None

    def _stage_inputs(self, args):
        import numpy as np
        from .core.tensor import Tensor
        if self.static_in is None:
            self.static_in = []
            for a in args:
                if not isinstance(a, Tensor) or not a.device.is_cuda:
                    raise TypeError("graphed_step: arguments must be cuda Tensors")
                with a.device:
                    self.static_in.append(Tensor(a.data, dtype=a.dtype, device=a.device, copy=True))
            return
        if len(args) != len(self.static_in):
            raise ValueError("graphed_step: number of arguments changed")
        for s, a in zip(self.static_in, args):
            if a is s:
                continue
            d = a.data if isinstance(a, Tensor) else np.asarray(a)
            if tuple(d.shape) != tuple(s.shape) or np.dtype(d.dtype) != s.dtype:
                raise ValueError(f"graphed_step: argument changed from {s.shape} {s.dtype} to {tuple(d.shape)} {d.dtype}")
            with s.device:
                s.data[...] = d

    def __call__(self, *args):
        self.calls += 1
        if self.calls <= self.warmup:
            return self.fn(*args)
        self._stage_inputs(args)
        dev = self.static_in[0].device if self.static_in else Device("cuda")
        flats = [o._flat for o in self.optimizers if getattr(o, "_flat", None) is not None]
        with dev:
            if self.graph is None:
                for o in self.optimizers:
                    if getattr(o, "_flat", None) is not None:
                        o._flat.sync_device_state(o.t, o.lr)
                g = Graph()
                g.begin()
                try:
                    self.out = self.fn(*self.static_in)
                finally:
                    g.end()
                self.graph = g
                for o in self.optimizers:  # the recording itself launched nothing: undo the host-side step count of fn
                    if getattr(o, "_flat", None) is not None:
                        o.t -= 1
            for o in self.optimizers:
                fl = getattr(o, "_flat", None)
                if fl is not None and fl.device_state()["lr_host"] != float(o.lr):
                    fl.sync_device_state(o.t, o.lr)
            self.graph.launch()
        for o in self.optimizers:
            if getattr(o, "_flat", None) is not None:
                o.t += 1
        for fl in flats:  # caches keyed on a buffer's write counter (operand planes, inference plans) must see the update
            fl.flat_p.buf.version += 1
            fl.flat_g.buf.version += 1
        return self.out
