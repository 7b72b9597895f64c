"""Datasets, samplers and the DataLoader of the reference (pydynet/data.py:4-123) plus the B200 input pipeline behind it.

With ``device=None`` a DataLoader behaves exactly like the reference's: every batch is ``dataset[list_of_indices]`` in the order its
samplers produce (``RandomSampler`` draws ONE ``numpy.random.permutation`` per epoch from the global NumPy stream, so seeded runs see
the reference's batches), and the training loop uploads it with a synchronous ``Tensor(batch)`` (examples/pydynet/mnist.py:161-162).

With ``device="cuda:i"`` the loader yields device Tensors instead: the NumPy arrays of batch i+1 are gathered into one of two PINNED
staging buffers and copied on a dedicated copy stream (``pdn_prefetch_h2d``) while the compute stream works on batch i; the consumer's
stream is ordered after the copy with an event, never with a host synchronisation."""
import ctypes as C

import numpy as np
from numpy.random import permutation


class Dataset:

    def __init__(self) -> None:
        pass

    def __getitem__(self, index):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError


class Sampler:

    def __init__(self, dataset: Dataset) -> None:
        pass

    def __iter__(self):
        raise NotImplementedError


class SequentialSampler(Sampler):

    def __init__(self, dataset: Dataset) -> None:
        self.dataset = dataset

    def __iter__(self):
        return iter(range(len(self.dataset)))

    def __len__(self) -> int:
        return len(self.dataset)


class RandomSampler(Sampler):

    def __init__(self, dataset: Dataset) -> None:
        self.dataset = dataset

    def __iter__(self):
        yield from permutation(len(self.dataset)).tolist()

    def __len__(self):
        return len(self.dataset)


class BatchSampler(Sampler):

    def __init__(self, sampler: Sampler, batch_size: int, drop_last: bool) -> None:
        self.sampler, self.batch_size, self.drop_last = sampler, batch_size, drop_last

    def __iter__(self):
        batch = []
        for idx in self.sampler:
            batch.append(idx)
            if len(batch) == self.batch_size:
                yield batch
                batch = []
        if batch and not self.drop_last:
            yield batch

    def __len__(self):
        n = len(self.sampler)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size


class _HostIter:
    """The reference's iterator: one ``dataset[indices]`` per step, nothing else."""

    def __init__(self, loader) -> None:
        self.loader = loader
        self.sample_iter = iter(loader.batch_sampler)

    def __iter__(self):
        return self

    def __next__(self):
        return self.loader.dataset[next(self.sample_iter)]


class _Staging:
    """One pinned host staging buffer + the event recorded after the last copy issued from it."""

    def __init__(self):
        self.ptr, self.size, self.event, self.used = C.c_void_p(), 0, C.c_void_p(), False

    def view(self, lib, nbytes):
        if self.used:
            lib.call("pdn_event_synchronize", self.event)  # the previous copy out of this buffer has finished
            self.used = False
        if nbytes > self.size:
            if self.ptr:
                lib.call("pdn_free_host", self.ptr)
            self.size = max(int(nbytes * 1.25), 1 << 16)
            lib.call("pdn_malloc_host", C.byref(self.ptr), self.size)
        if not self.event:
            lib.call("pdn_event_create", C.byref(self.event))
        return np.frombuffer((C.c_byte * self.size).from_address(self.ptr.value), dtype=np.uint8)

    def release(self, lib):
        try:
            if self.used:
                lib.call("pdn_event_synchronize", self.event)
            if self.ptr:
                lib.call("pdn_free_host", self.ptr)
            if self.event:
                lib.call("pdn_event_destroy", self.event)
        except Exception:
            pass
        self.ptr, self.size, self.event, self.used = C.c_void_p(), 0, C.c_void_p(), False


class _PrefetchIter:
    """Device iterator: batch i+1 is in flight on the copy stream while batch i is consumed."""

    def __init__(self, loader) -> None:
        from .backend import lib
        from .cuda import Device
        self.loader, self.lib, self.dev = loader, lib, Device(loader.device)
        self.sample_iter = iter(loader.batch_sampler)
        self.slots, self.turn, self.pending = [_Staging(), _Staging()], 0, None
        self._issue()

    def __iter__(self):
        return self

    def _issue(self):
        try:
            index = next(self.sample_iter)
        except StopIteration:
            self.pending = None
            return
        from .backend.array import ndarray
        from .core.tensor import Tensor
        items = self.loader.dataset[index]
        single = not isinstance(items, (tuple, list))
        seq = [items] if single else list(items)
        host = []
        for it in seq:
            if isinstance(it, Tensor) and not it.device.is_cuda:  # the examples keep the whole set in cpu Tensors (mnist.py:143-152)
                it = np.asarray(it.data)
            if isinstance(it, np.ndarray) and it.dtype != object:
                a = np.ascontiguousarray(it if self.loader.dtype is None or it.dtype.kind != "f" else it.astype(self.loader.dtype, copy=False))
                host.append(a)
            else:
                host.append(None)
        offs, total = [], 0
        for a in host:
            offs.append(total)
            if a is not None:
                total += (a.nbytes + 255) // 256 * 256
        slot = self.slots[self.turn]
        self.turn ^= 1
        with self.dev:
            stage = slot.view(self.lib, total)
            out = []
            for it, a, off in zip(seq, host, offs):
                if a is None:
                    out.append(it)
                    continue
                stage[off:off + a.nbytes] = a.reshape(-1).view(np.uint8)
                d = ndarray.empty(a.shape, a.dtype)
                self.lib.call("pdn_prefetch_h2d", d.ptr, C.c_void_p(slot.ptr.value + off), a.nbytes, slot.event)
                slot.used = True
                out.append(d)
        self.pending = (slot, out, single)

    def __next__(self):
        if self.pending is None:
            for s in self.slots:
                s.release(self.lib)
            raise StopIteration
        from .core.tensor import Tensor
        from .backend.array import ndarray
        slot, out, single = self.pending
        with self.dev:
            if slot.used:
                self.lib.call("pdn_stream_wait_event", slot.event)  # compute stream after the copy; the host does not wait
            res = [Tensor(o, dtype=o.dtype, copy=None, device=self.dev) if isinstance(o, ndarray) else o for o in out]
        self._issue()  # the next batch's gather + copy overlap with whatever the caller launches on this one
        return res[0] if single else tuple(res)

    def __del__(self):
        for s in getattr(self, "slots", []):
            s.release(self.lib)


class DataLoader:

    def __init__(self, dataset: Dataset, batch_size: int = 1, shuffle: bool = False, drop_last: bool = False, device=None, dtype=None) -> None:
        self.dataset, self.batch_size, self.shuffle, self.drop_last = dataset, batch_size, shuffle, drop_last
        self.sampler = RandomSampler(dataset) if shuffle else SequentialSampler(dataset)
        self.batch_sampler = BatchSampler(self.sampler, batch_size, drop_last)
        self.device = None if device is None or str(device) == "cpu" else device
        self.dtype = dtype  # optional floating dtype the NumPy batches are cast to on the host before the upload

    def __iter__(self):
        return _HostIter(self) if self.device is None else _PrefetchIter(self)

    def __len__(self):
        return len(self.batch_sampler)


def data_loader(X, y, batch_size: int, shuffle: bool = False, device=None, dtype=None) -> DataLoader:

    class TrainSet(Dataset):

        def __init__(self, X, y) -> None:
            self.data, self.target = X, y

        def __getitem__(self, index):
            return self.data[index], self.target[index]

        def __len__(self):
            return len(self.data)

    return DataLoader(TrainSet(X, y), batch_size, shuffle, device=device, dtype=dtype)
