"""Boundary variant B: the reference's OWN engine — its unmodified pydynet/core/tensor.py, function.py,
autograd.py, cuda.py — running on ``pydynet_b200.backend`` as its array module, in the slot CuPy has in the reference
(cuda.py:4-13, 90-91: ``xp = np if cpu else cp``).  The reference package is imported from the staged tree with a module named
``cupy`` that is this repo's backend (+ the three CuPy runtime calls cuda.py makes), its default device is switched to cuda:0, and
its own 77-case test-suite (NumPy equality of every elementwise / matmul / reduction / shape operator incl. dtype promotion, and the
backward cases) must pass: every array expression of the reference's operators then executes in libpdn_b200.so kernels.
Runs in a subprocess (module aliasing must not leak)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline.stage_reference import reference_root  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(reference_root() is None or not os.path.isdir(os.path.join(reference_root(), "tests")),
                                                  reason="reference tree (with its tests) neither mounted nor staged")]

DRIVER = r'''
import sys, types
sys.path.insert(0, %(root)r)
import numpy as np
import pydynet_b200.backend as b200
import pydynet_b200.cuda as our_cuda
from pydynet_b200.backend import lib as L

# ---- a module named `cupy`: the array-module surface is pydynet_b200.backend, plus what reference cuda.py asks of CuPy's runtime
cp = types.ModuleType("cupy")
class _Shim(types.ModuleType):
    def __getattr__(self, name):
        return getattr(b200, name)
cp.__class__ = _Shim
class _Dev:
    def __init__(self, i): self.id = int(i)
    def __enter__(self):
        self._prev = our_cuda.current_device(); our_cuda.set_device(self.id); return self
    def __exit__(self, *a): our_cuda.set_device(self._prev)
    def __eq__(self, o): return isinstance(o, _Dev) and o.id == self.id
    def __hash__(self): return hash(self.id)
runtime = types.SimpleNamespace(getDeviceCount=our_cuda.device_count, getDevice=our_cuda.current_device, setDevice=our_cuda.set_device)
cp.cuda = types.SimpleNamespace(runtime=runtime, Device=_Dev)
cp.add = types.SimpleNamespace(at=b200.add_at)
sys.modules["cupy"] = cp
our_cuda.Device("cuda:0")  # creates the context / streams

sys.path.insert(0, %(ref)r)
import pydynet                      # the UNMODIFIED reference package
assert pydynet.__file__.startswith(%(ref)r), pydynet.__file__
assert pydynet.cuda.is_available()
# default device of the suite's tensors: cuda:0 (the suite builds Tensor(ndarray) without a device argument)
_orig = pydynet.cuda.Device.__init__
def _init(self, device=None):
    _orig(self, "cuda:0" if device is None else device)
pydynet.cuda.Device.__init__ = _init
t = pydynet.Tensor(np.arange(6.).reshape(2, 3))
assert type(t.data).__module__.startswith("pydynet_b200"), type(t.data)
L.reset_launch_count()
import pytest
rc = pytest.main(["-q", "-p", "no:cacheprovider", "--tb=line", %(tests)r])
print("KERNEL_LAUNCHES", L.launch_count())
sys.exit(rc)
'''


def test_reference_engine_on_b200_backend_passes_its_own_suite(tmp_path):
    ref = reference_root()
    # the suite hands raw device arrays (``x.grad``) to np.testing / np.allclose: like CuPy, the backend refuses implicit host
    # conversion unless asked to
    r = subprocess.run([sys.executable, "-c", DRIVER % {"root": ROOT, "ref": ref, "tests": os.path.join(ref, "tests")}], capture_output=True,
                       text=True, cwd=str(tmp_path), timeout=900, env=dict(os.environ, PDN_IMPLICIT_NUMPY="1"))
    tail = (r.stdout + r.stderr)[-6000:]
    assert r.returncode == 0, tail
    assert "77 passed" in r.stdout, tail
    launches = int(r.stdout.split("KERNEL_LAUNCHES")[-1].split()[0])
    assert launches > 100, launches  # the suite really ran on the device (205 kernel launches)
