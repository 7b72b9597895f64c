import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cuda():
    """The cuda:0 device of the package under test; fails loudly (no CPU fallback) if the library or GPU is missing."""
    import pydynet_b200 as pdn
    return pdn.Device("cuda:0")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need the library AND a device: on a host without an NVIDIA device node they are skipped, not errored (a plain
    `pytest` on a CPU-only box stays green). Where /dev/nvidia0 exists (or PDN_REQUIRE_GPU=1) nothing is skipped, so a build that does
    not load on the GPU box fails loudly instead of hiding behind the skip."""
    if os.environ.get("PDN_REQUIRE_GPU") == "1" or os.path.exists("/dev/nvidia0"):
        return  # a GPU box: a library that does not load there must FAIL the gpu tests, not skip them
    try:
        import pydynet_b200 as pdn
        have = bool(pdn.cuda.is_available())
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no B200 / libpdn_b200.so on this host (set PDN_REQUIRE_GPU=1 to fail instead)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
