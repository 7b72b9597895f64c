"""CPU-only checks of the drop-in boundary: libpdn_b200.so loads, exports every symbol declared in include/pdn_b200.h
(and only binds what the header declares), and the cuda device fails loudly — never silently on the CPU — when no GPU is
present."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pdn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pdn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pydynet_b200.backend import lib
    l = lib.load()
    assert lib.MISSING == [], f"declared in lib.py but not exported: {lib.MISSING}"
    names = _declared()
    assert len(names) > 60
    for n in names:
        assert hasattr(l, n), f"include/pdn_b200.h declares {n} but libpdn_b200.so does not export it"
    bound = set(lib.declared_symbols())
    assert set(names) == bound, f"header vs ctypes binding mismatch: {sorted(set(names) ^ bound)}"


def test_header_cites_reference_call_sites():
    src = open(os.path.join(ROOT, "include", "pdn_b200.h")).read()
    assert len(re.findall(r"[a-z_/]+\.py:\d+", src)) >= 30


def test_cuda_device_fails_loudly_without_gpu():
    import pydynet_b200 as pdn
    if pdn.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError):
        pdn.Device("cuda:0")
    with pytest.raises(RuntimeError):
        pdn.Tensor([1.0], device="cuda")


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pydynet_b200")):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text or "import oracle" not in text and "from oracle" not in text, f
