"""Runs tests/differential_vs_reference.py (this package vs the live reference, cpu device) in a subprocess; skipped where the
reference is not mounted."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/pydynet"), reason="reference sources not mounted")


def test_modules_and_optimizers_match_live_reference():
    r = subprocess.run([sys.executable, os.path.join(HERE, "differential_vs_reference.py")], capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert r.returncode == 0 and "bad: 0" in r.stdout, (r.stdout + r.stderr)[-2000:]
    assert r.stdout.count(": ok") >= 19
