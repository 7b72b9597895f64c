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
    """Rank 0 creates the 128-byte NCCL id; everybody else receives it (torch.distributed if it is up, else a file)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            box = [make_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            return box[0]
    except ImportError:
        pass
    tag = f"{os.environ.get('MASTER_PORT', '0')}_{os.environ.get('TORCHELASTIC_RUN_ID', os.environ.get('PDN_JOB_ID', 'job'))}"
    path = os.path.join(os.environ.get("PDN_STORE_DIR", "/tmp"), f"pdn_nccl_id_{tag}")
    if rank == 0:
        data = make_id()
        with open(path + ".tmp", "wb") as f:
            f.write(data)
        os.replace(path + ".tmp", path)
        return data
    deadline = time.time() + 120
    while not os.path.exists(path):
        if time.time() > deadline:
            raise RuntimeError(f"timed out waiting for the NCCL id at {path}")
        time.sleep(0.01)
    with open(path, "rb") as f:
        return f.read()


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
    """

    def __init__(self, module, optimizer):
        self.module, self.optimizer = module, optimizer
        self.world = _S["world"]
        self._flat = getattr(optimizer, "_flat", None)
        if self._flat is not None:
            optimizer.grad_scale = 1.0 / self.world  # folded into the fused Adam kernel

    def __call__(self, *a):
        return self.module(*a)

    def zero_grad(self):
        self.optimizer.zero_grad()

    def sync_gradients(self):
        if self.world == 1:
            return
        if self._flat is not None:
            from .backend import lib
            with self._flat.device:
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
