"""Fused autograd nodes for fp32 cuda tensors: one hand-written kernel (or a short fixed sequence) per functional op,
each a single tape entry.  Definitions of the results are the operator chains in nn/functional.py and nn/modules/*;
this module only changes how many launches and HBM passes they cost.  There is no host path in here: every function
raises if the backend library is missing."""
import ctypes as C
import os

import numpy as np

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


def sequence(cell, x, state):
    """Fused whole-sequence recurrence for one layer/direction; None = not applicable, caller runs the per-step loop."""
    return None
