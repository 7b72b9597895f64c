"""The tensor-core kernels must really be tcgen05 / TMEM / TMA code: checks the SASS of the built libpdn_b200.so for the Blackwell
mnemonics (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tile loads, UTCBAR = tcgen05.commit) in every tensor
kernel family and for the absence of the legacy mma.sync path (HMMA). CPU-only: cuobjdump reads the cross-compiled library."""
import os
import shutil
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "pydynet_b200", "libpdn_b200.so")
pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(SO), reason="needs cuobjdump and the built library")


def test_tensor_kernels_use_tcgen05_tmem_tma():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from sass_evidence import evidence
    ev = evidence(SO)
    fam = {"k_gemm_tc": [], "k_conv_tma": [], "k_attn_tc": []}
    for name, c in ev.items():
        for f in fam:
            if f + "<" in name:
                fam[f].append((name, c))
    for f, kernels in fam.items():
        assert kernels, f"no {f} instantiation in the library"
        for name, c in kernels:
            assert c.get("UTCHMMA", 0) > 0 and c.get("UTMALDG", 0) > 0 and c.get("LDTM", 0) > 0 and c.get("UTCBAR", 0) > 0, (name, c)
            assert c.get("ELECT", 0) > 0, (name, c)  # warp-uniform issue: elect.sync, not `if (lane == 0)`
    # flash attention keeps P / dS and the row operand in tensor memory (tcgen05.st)
    assert all(c.get("STTM", 0) > 0 for _, c in fam["k_attn_tc"])
    # nothing falls back to the legacy warp-level tensor path
    assert not any(c.get("HMMA", 0) or c.get("HGMMA", 0) for c in ev.values())
