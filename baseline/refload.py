"""Test / measurement infrastructure — NOT product code.

Loads the UNMODIFIED reference's files (from /root/reference, or the staged copy under baseline/_ref on the GPU box —
see baseline/stage_reference.py) in two ways:

* ``ref_model(path)``      exec's a reference model / example file against the REAL reference package (NumPy path): the live
                           oracle and the ``--impl reference`` arm of the benches;
* ``dropin_model(path)``   exec's the SAME file with ``pydynet`` aliased to ``pydynet_b200``: the drop-in check — the
                           reference's own model classes running on this backend without edits.
"""
import contextlib
import importlib
import os
import sys

from .stage_reference import reference_root

_ALIASES = ("pydynet", "pydynet.core", "pydynet.core.tensor", "pydynet.core.function", "pydynet.nn", "pydynet.nn.functional",
            "pydynet.nn.parameter", "pydynet.nn.init", "pydynet.nn.modules", "pydynet.special", "pydynet.optim", "pydynet.autograd",
            "pydynet.cuda", "pydynet.data")


def available() -> bool:
    return reference_root() is not None


@contextlib.contextmanager
def _modules(mapping):
    saved = {k: sys.modules.get(k) for k in mapping}
    # anything else below 'pydynet.' that a previous context imported must not leak across
    stale = {k: sys.modules.pop(k) for k in list(sys.modules) if (k == "pydynet" or k.startswith("pydynet.")) and k not in mapping}
    sys.modules.update(mapping)
    try:
        yield
    finally:
        for k in list(sys.modules):
            if k == "pydynet" or k.startswith("pydynet."):
                sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
        sys.modules.update(stale)


def dropin_aliases():
    """``pydynet*`` module names -> pydynet_b200 modules (what a user gets by `import pydynet_b200 as pydynet`)."""
    out = {}
    for name in _ALIASES:
        ours = "pydynet_b200" + name[len("pydynet"):]
        try:
            out[name] = importlib.import_module(ours)
        except ImportError:
            pass
    return out


def aliased():
    """Context: `import pydynet` resolves to pydynet_b200."""
    return _modules(dropin_aliases())


_REF_PKG = {}


def reference_package():
    """The real reference package object (imported once from the reference root, kept OUT of sys.modules afterwards)."""
    if "pkg" not in _REF_PKG:
        root = reference_root()
        if root is None:
            raise RuntimeError("reference tree not available (neither /root/reference nor baseline/_ref)")
        import warnings
        with _modules({}):
            sys.path.insert(0, root)
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    pkg = importlib.import_module("pydynet")
                    for sub in ("pydynet.nn", "pydynet.nn.functional", "pydynet.optim", "pydynet.core.tensor", "pydynet.special",
                                "pydynet.nn.parameter", "pydynet.autograd", "pydynet.cuda", "pydynet.nn.init", "pydynet.data"):
                        try:
                            importlib.import_module(sub)
                        except ImportError:
                            pass
                _REF_PKG["mods"] = {k: v for k, v in sys.modules.items() if k == "pydynet" or k.startswith("pydynet.")}
            finally:
                sys.path.remove(root)
        _REF_PKG["pkg"] = pkg
    return _REF_PKG["pkg"]


def referenced():
    """Context: `import pydynet` resolves to the real reference."""
    reference_package()
    return _modules(dict(_REF_PKG["mods"]))


def _exec_file(rel, lines=None, extra=None, name="ref_file"):
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not available (neither /root/reference nor baseline/_ref)")
    path = os.path.join(root, rel)
    src = open(path).read().splitlines()
    if lines is not None:
        src = src[lines[0]:lines[1]]
    ns = {"__name__": name}
    ns.update(extra or {})
    exec(compile("\n".join(src), path, "exec"), ns)
    return ns


def dropin_model(rel, lines=None, extra=None):
    """Namespace of reference file ``rel`` (optionally a line window) exec'd on top of pydynet_b200."""
    with aliased():
        return _exec_file(rel, lines, extra, "ref_file_on_b200")


def ref_model(rel, lines=None, extra=None):
    """Namespace of reference file ``rel`` exec'd on top of the real reference package."""
    with referenced():
        return _exec_file(rel, lines, extra, "ref_file_on_numpy")


def dropin_extra():
    import numpy as np
    import pydynet_b200 as pdn
    return {"np": np, "pdn": pdn, "nn": pdn.nn, "F": pdn.nn.functional, "DTYPE": np.float32}


def ref_extra():
    import numpy as np
    pkg = reference_package()
    m = _REF_PKG["mods"]
    return {"np": np, "pdn": pkg, "nn": m["pydynet.nn"], "F": m["pydynet.nn.functional"], "DTYPE": np.float32}
