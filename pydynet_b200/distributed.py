"""Data-parallel training: one process per GPU, batch sharded across ranks, parameters replicated, ONE exchange step per
iteration — a sum all-reduce of the flat fp32 gradient bucket (NCCL over NVLink/NVSwitch, csrc/comm.cu) whose 1/world
scale is folded into the fused Adam kernel.

The reference is single-process (no NCCL / MPI / distributed code); this module is the exchange step
BASELINE's north_star defines. Parity contract: with equal shards, W ranks produce the parameters a single
process produces on the concatenated batch. That needs, besides gradient averaging, GLOBAL batch statistics in the
batch-coupled norms (BatchNorm*, the reference's "LayerNorm"): `sync_batch_stats` makes their forward / backward reduce
per-feature sums across ranks (2 x C floats each way).

Backends: "nccl" — cuda tensors, libpdn_b200's communicator; "gloo" — cpu-device tensors through torch.distributed
(used by the world_size-2 CPU tests of this host logic).
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

from .core.tensor import Tensor, _result

_S = {"backend": None, "rank": 0, "world": 1, "sync_stats": False}


def is_initialized() -> bool:
    return _S["backend"] is not None


def get_rank() -> int:
    return _S["rank"]


def get_world_size() -> int:
    return _S["world"]


def sync_batch_stats(enable: bool = True) -> None:
    _S["sync_stats"] = bool(enable)


def sync_stats_enabled() -> bool:
    return _S["sync_stats"] and _S["world"] > 1


def _exchange_id(rank, world, make_id) -> bytes:
    """Rank 0 creates the 128-byte NCCL id; everybody else receives it: through torch.distributed when a process group is up,
    else over a one-shot TCP rendezvous on MASTER_ADDR:(MASTER_PORT + 23) (``PDN_ID_PORT`` overrides the port). Nothing is left
    behind on disk, so a second job on the same host and port can never pick up a previous job's id."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            box = [make_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            return box[0]
    except ImportError:
        pass
    import socket
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(os.environ.get("PDN_ID_PORT", int(os.environ.get("MASTER_PORT", "29500")) + 23))
    if rank == 0:
        data = make_id()
        with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as srv:
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, port))
            srv.listen(world)
            srv.settimeout(120)
            for _ in range(world - 1):
                conn, _peer = srv.accept()
                with conn:
                    conn.sendall(data)
        return data
    deadline = time.time() + 120
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5) as c:
                data = b""
                while len(data) < 128:
                    chunk = c.recv(128 - len(data))
                    if not chunk:
                        break
                    data += chunk
            if len(data) == 128:
                return data
        except OSError:
            pass
        if time.time() > deadline:
            raise RuntimeError(f"timed out waiting for the NCCL id from rank 0 at {addr}:{port}")
        time.sleep(0.05)


def init_process_group(backend: str = "nccl", rank: int | None = None, world_size: int | None = None) -> None:
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", 1)) if world_size is None else world_size
    if backend == "nccl":
        from .backend import lib
        from . import cuda
        local = int(os.environ.get("LOCAL_RANK", rank))
        cuda.Device(f"cuda:{local}")
        cuda.set_device(local)

        def make_id():
            buf = C.create_string_buffer(128)
            lib.call("pdn_nccl_unique_id", buf)
            return buf.raw

        uid = _exchange_id(rank, world, make_id)
        lib.call("pdn_nccl_init", rank, world, C.create_string_buffer(uid, 128))
    elif backend == "gloo":
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        raise ValueError(f"unknown backend {backend!r}")
    _S.update(backend=backend, rank=rank, world=world)


def destroy_process_group() -> None:
    if _S["backend"] == "nccl":
        from .backend import lib
        lib.call("pdn_nccl_destroy")
    _S.update(backend=None, rank=0, world=1, sync_stats=False)


def all_reduce_sum_(arr):
    """In-place sum across ranks of a NumPy array (gloo) or a contiguous fp32 device array (nccl, on the compute stream)."""
    if _S["world"] == 1:
        return arr
    if isinstance(arr, np.ndarray):
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(arr))
        dist.all_reduce(t)
        if t.numpy() is not arr:
            arr[...] = t.numpy()
        return arr
    from .backend import lib
    assert arr.dtype == np.float32 and arr.is_contiguous
    lib.call("pdn_allreduce_sum_f32_inline", arr.ptr, arr.size)
    arr.buf.version += 1
    return arr


def dist_mean(t: Tensor) -> Tensor:
    """Mean over ranks as an autograd node: forward all_reduce(t)/W, backward all_reduce(g)/W — the gradient of every
    rank's loss with respect to a shared statistic flows back to every rank's contribution."""
    W = _S["world"]
    if W == 1:
        return t
    xp = t.xp
    with t.device:
        data = all_reduce_sum_(xp.array(t.data, copy=True)) / W

    def backward(g):
        return (all_reduce_sum_(xp.array(g, dtype=g.dtype, copy=True)) / W, )

    return _result(data, t.device, (t, ), backward, "dist_mean")


def shard(array, axis: int = 0):
    """This rank's equal slice of a global batch along ``axis``."""
    n = array.shape[axis]
    W, r = _S["world"], _S["rank"]
    assert n % W == 0, "the global batch must divide evenly across ranks"
    sl = [slice(None)] * array.ndim
    sl[axis] = slice(r * n // W, (r + 1) * n // W)
    return array[tuple(sl)]


class DataParallel:
    """Wraps (module, optimizer): ``step()`` = all-reduce gradients, then ``optimizer.step()`` with the 1/world scale.

        ddp = DataParallel(net, Adam(net.parameters()))
        loss = loss_fn(net(shard(X)), shard(y)); ddp.zero_grad(); loss.backward(); ddp.step()

    With the flat Adam state (cuda, fp32) the gradient bucket is cut into ``buckets`` contiguous ranges of the flat gradient buffer
    and, with ``overlap=True``, the backward sweep reports every parameter whose gradient is final (core/tensor.py,
    ``_LEAF_READY_HOOK``): as soon as all parameters of a range are final its ``ncclAllReduce`` is queued on the side stream
    (ordered after the compute stream by an event) while the rest of the backward pass keeps computing; ``step()`` queues the
    ranges that are left (parameters without a gradient this step) and makes the compute stream wait for the side stream before
    the fused Adam kernel. One backward per step (gradient accumulation across several backward calls needs ``overlap=False``).
    Construction makes the replicas identical: parameters, Adam moments and registered buffers of rank 0 reach every rank."""

    def __init__(self, module, optimizer, buckets: int = 4, overlap: bool = True, broadcast: bool = True):
        self.module, self.optimizer = module, optimizer
        self.world = _S["world"]
        self._flat = getattr(optimizer, "_flat", None)
        self.overlap = bool(overlap and self._flat is not None and self.world > 1 and _S["backend"] == "nccl")
        self._ranges, self._owner, self._left, self._launched, self._armed = [], {}, [], [], False
        self.launch_log = []  # (bucket index, number of tape sweeps finished when it was queued): evidence of the overlap
        self._sweeps = 0
        if self._flat is not None:
            optimizer.grad_scale = 1.0 / self.world  # folded into the fused Adam kernel
            self._make_buckets(max(1, int(buckets)))
        if broadcast and self.world > 1:
            self._broadcast_state()
        self._arm()

    # -------------------------------------------------------------------------------------------------- set-up
    def _make_buckets(self, n):
        fl = self._flat
        sizes = [int(o2 - o1) for o1, o2 in zip(fl.offs, fl.offs[1:] + [fl.total])]
        target, ranges, start, acc, members = fl.total / n, [], 0, 0, []
        for i, (o, sz) in enumerate(zip(fl.offs, sizes)):
            members.append(i)
            acc += sz
            if acc >= target and len(ranges) < n - 1:
                ranges.append((start, o + sz - start, members))
                start, acc, members = o + sz, 0, []
        if members:
            ranges.append((start, fl.total - start, members))
        self._ranges = ranges
        self._owner = {id(fl.params[i]): b for b, (_, _, mem) in enumerate(ranges) for i in mem}

    def _broadcast_state(self):
        """Rank 0's parameters / Adam moments / running statistics everywhere: the other ranks zero theirs and a sum all-reduce
        delivers rank 0's values exactly (x + 0 == x)."""
        arrays = []
        if self._flat is not None:
            arrays += [self._flat.flat_p, self._flat.flat_m, self._flat.flat_v]
        else:
            arrays += [p.data for p in self.optimizer.params]
        for m in self._modules(self.module):
            for nm in ("running_mean", "running_var"):
                t = getattr(m, nm, None)
                if isinstance(t, Tensor):
                    arrays.append(t.data)
        for a in arrays:
            dev = self._flat.device if self._flat is not None else None
            if isinstance(a, np.ndarray):
                if _S["rank"] != 0:
                    a[...] = 0
                all_reduce_sum_(a)
            elif a.dtype == np.float32 and a.is_contiguous:
                with (dev or self.optimizer.params[0].device):
                    if _S["rank"] != 0:
                        a.fill(0.0)
                    all_reduce_sum_(a)

    @staticmethod
    def _modules(root):
        out, stack = [], [root]
        while stack:
            m = stack.pop()
            out.append(m)
            for v in m.__dict__.values():
                if hasattr(v, "_parameters") and hasattr(v, "forward"):
                    stack.append(v)
                elif isinstance(v, (list, tuple)):
                    stack.extend(x for x in v if hasattr(x, "_parameters") and hasattr(x, "forward"))
        return out

    def __call__(self, *a):
        return self.module(*a)

    # -------------------------------------------------------------------------------------------------- per step
    def zero_grad(self):
        self.optimizer.zero_grad()

    def _arm(self):
        """The NEXT backward sweep reports finished parameters (armed at construction and after every ``step()``, so the usual loop
        ``opt.zero_grad(); loss.backward(); ddp.step()`` overlaps without further calls)."""
        if self.overlap:
            self._flat.repin()
            self._left = [len(mem) for (_, _, mem) in self._ranges]
            self._launched = [False] * len(self._ranges)
            self._sweeps = 0
            self._armed = True
            from .core import tensor as _t
            _t._LEAF_READY_HOOK[0] = self._leaf_ready

    def _leaf_ready(self, leaf):
        if leaf is None:  # a backward sweep finished
            self._sweeps += 1
            from .core import tensor as _t
            _t._LEAF_READY_HOOK[0] = None  # one sweep per step feeds the early launches
            return
        b = self._owner.get(id(leaf))
        if b is None or not self._armed:
            return
        self._left[b] -= 1
        if self._left[b] == 0:
            self._launch(b)

    def _launch(self, b):
        from .backend import lib
        fl = self._flat
        off, n, mem = self._ranges[b]
        with fl.device:
            for i in mem:  # a parameter of this range that received no gradient holds stale values: its gradient is zero
                p = fl.params[i]
                if p._grad_stale:
                    p._grad.fill(0.0)
                    p._grad_stale = False
            fl.flat_g.buf.version += 1
            lib.call("pdn_allreduce_sum_f32", fl.flat_g.ptr + 4 * off, n)  # side stream, after what the compute stream queued so far
        self._launched[b] = True
        if len(self.launch_log) < 64:
            self.launch_log.append((b, self._sweeps))

    def sync_gradients(self):
        if self.world == 1:
            return
        from .cuda import nvtx_range
        with nvtx_range("grad_allreduce"):
            self._sync_gradients()

    def _sync_gradients(self):
        if self._flat is not None:
            from .backend import lib
            if self.overlap and self._armed:
                if self._sweeps > 1:
                    raise RuntimeError("DataParallel(overlap=True) supports one backward per step; use overlap=False to accumulate")
                for b in range(len(self._ranges)):
                    if not self._launched[b]:
                        self._launch(b)
                self._armed = False
                from .core import tensor as _t
                _t._LEAF_READY_HOOK[0] = None
                with self._flat.device:
                    lib.call("pdn_allreduce_wait")
                return
            with self._flat.device:
                self._flat.repin()
                self._flat._settle_grads()
                self._flat.flat_g.buf.version += 1
                lib.call("pdn_allreduce_sum_f32", self._flat.flat_g.ptr, self._flat.total)  # comm stream, after backward
                lib.call("pdn_allreduce_wait")  # compute stream resumes when the bucket is reduced
            return
        for p in self.optimizer.params:
            g = p.grad
            with p.device:
                if isinstance(g, np.ndarray):
                    all_reduce_sum_(g)
                    g /= self.world
                else:
                    gc = g if (g.is_contiguous and g.dtype == np.float32) else g.astype(np.float32).copy()
                    all_reduce_sum_(gc)
                    gc /= self.world
                    if gc is not g:
                        g[...] = gc

    def step(self):
        self.sync_gradients()
        self.optimizer.step()
        self._arm()
