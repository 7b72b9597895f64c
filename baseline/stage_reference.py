"""Test / measurement infrastructure — NOT product code (nothing under pydynet_b200/ imports this; lives beside baseline/_ref).

Stages an UNMODIFIED copy of the reference (WeltXing/PyDyNet, /root/reference) into ``baseline/_ref/`` so that it travels
to the GPU box with the gpurun snapshot (``baseline/_ref/`` is git-ignored, not gpurun-ignored: no reference source ever
enters this repository's history).  It is used for exactly two things:

* ``bench.py --impl reference`` / ``bench_all.py``: the reference's own NumPy path timed on the box's host cores
  (``cpu_baseline.kind == "reference"``);
* ``tests/``: the reference's own model files (``llm/llama/model.py``, ``examples/pydynet/*.py``) exec'd with ``pydynet``
  aliased to ``pydynet_b200`` — the drop-in check on ``cuda`` — and the reference's array module as the live oracle.

``stage()`` follows the base contract's install recipe (``pip install --no-index --no-build-isolation --no-deps --target
baseline/_ref <copy of /root/reference>``; the copy is needed because the build writes egg-info into the source tree and
/root/reference is read-only) and then adds the model / example scripts, which ``setup.py`` does not package
(reference setup.py:12-15 lists only pydynet, pydynet/optim, pydynet/nn, pydynet/nn/modules, pydynet/core).
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
SCRIPT_DIRS = ("llm/llama", "llm/clip", "examples/pydynet", "tests")  # python files only (no weights / images)


def reference_root():
    """Directory that holds the unmodified reference tree: /root/reference where it is mounted (the build container), else the
    staged copy that travelled with the snapshot, else None."""
    if os.environ.get("PDN_NO_REFERENCE") == "1":  # tests of the fallback paths (stand-in model definitions, oracle port as the CPU arm)
        return None
    if os.path.isfile(os.path.join(SRC, "pydynet", "__init__.py")):
        return SRC
    if os.path.isfile(os.path.join(DST, "pydynet", "__init__.py")):
        return DST
    return None


def stage(verbose=False) -> str:
    """Returns 'pip', 'copy' (pip failed, plain copy of the package used) or 'absent' (no /root/reference here)."""
    if not os.path.isdir(SRC):
        return "absent"
    how = "pip"
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns("imgs", "*.png", "*.jpg", "*.npz", ".git"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
               "--target", DST, work]
        out = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            print(out.stdout[-2000:], out.stderr[-2000:])
        if out.returncode != 0 or not os.path.isfile(os.path.join(DST, "pydynet", "__init__.py")):
            how = "copy"
            shutil.copytree(os.path.join(SRC, "pydynet"), os.path.join(DST, "pydynet"), dirs_exist_ok=True,
                            ignore=shutil.ignore_patterns("__pycache__"))
    for d in SCRIPT_DIRS:
        src = os.path.join(SRC, d)
        if not os.path.isdir(src):
            continue
        dst = os.path.join(DST, d)
        os.makedirs(dst, exist_ok=True)
        for fn in os.listdir(src):
            if fn.endswith(".py"):
                shutil.copy2(os.path.join(src, fn), os.path.join(dst, fn))
    with open(os.path.join(DST, "STAGED_FROM"), "w") as f:
        f.write(f"{SRC} via {how}\n")
    return how


if __name__ == "__main__":
    print(stage(verbose=True))
