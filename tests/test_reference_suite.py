"""The reference's OWN test-suite (/root/reference/tests: 77 cases pinning NumPy-equality of the elementwise / matmul / reduction /
shape operators incl. dtype promotion, and a handful of backward cases — SURVEY.md §4) run UNCHANGED against this package:
``pydynet`` is aliased to ``pydynet_b200`` before the test modules import it. Runs in a subprocess (module aliasing must not leak),
skipped where the reference is not mounted (the GPU box)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="reference sources not mounted")

DRIVER = r'''
import importlib, sys
sys.path.insert(0, %r)
import pydynet_b200 as pdn
sys.modules.update({"pydynet": pdn, "pydynet.core": pdn.core, "pydynet.core.tensor": importlib.import_module("pydynet_b200.core.tensor"),
                    "pydynet.nn": pdn.nn, "pydynet.nn.functional": pdn.nn.functional, "pydynet.special": pdn.special,
                    "pydynet.optim": pdn.optim, "pydynet.autograd": pdn.autograd, "pydynet.cuda": pdn.cuda})
import pytest
sys.exit(pytest.main(["-q", "-p", "no:cacheprovider", %r]))
'''


def test_reference_test_suite_passes_on_this_package(tmp_path):
    r = subprocess.run([sys.executable, "-c", DRIVER % (ROOT, os.path.join(REF, "tests"))], capture_output=True, text=True, cwd=str(tmp_path),
                       timeout=600)
    tail = (r.stdout + r.stderr)[-1500:]
    assert r.returncode == 0, tail
    assert "77 passed" in r.stdout, tail
