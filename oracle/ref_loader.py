"""Import the *real* reference modules (read-only tree) under private aliases -- TEST INFRASTRUCTURE ONLY.

Used by ``oracle/make_golden.py`` in the build container (where ``/root/reference`` exists) to pin
``oracle/restate.py`` and to freeze golden vectors.  Nothing on the GPU box needs this file to succeed:
``find_reference()`` returns ``None`` there and callers skip.

The reference does not import as shipped: ``trainer/__init__.py`` is broken and
``matplotlib`` / ``visdom`` are absent, so the modules are exec'd one by one with stub dependencies and
registered under ``_ref_*`` names to avoid colliding with this repo's own ``Model`` / ``trainer`` packages.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

CANDIDATES = ("/root/reference", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref"))


def find_reference():
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "Model", "CycleGan.py")):
            return c
    return None


def _stub(name, **attrs):
    if name not in sys.modules:
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    return sys.modules[name]


def _exec(alias, path, package=None):
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    if package:
        mod.__package__ = package
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    """Returns a namespace with CycleGan, HdGan, layers, reg, transformer, utils modules of the reference."""
    root = find_reference()
    if root is None:
        raise FileNotFoundError("reference tree not found (looked in %s)" % (CANDIDATES,))
    plt = _stub("matplotlib.pyplot")
    _stub("matplotlib", pyplot=plt)
    _stub("visdom", Visdom=object)
    pkg = types.ModuleType("_ref_trainer")
    pkg.__path__ = [os.path.join(root, "trainer")]
    sys.modules["_ref_trainer"] = pkg
    ns = types.SimpleNamespace(root=root)
    ns.CycleGan = _exec("_ref_Model_CycleGan", os.path.join(root, "Model", "CycleGan.py"))
    ns.HdGan = _exec("_ref_Model_HdGan", os.path.join(root, "Model", "HdGan.py"))
    ns.layers = _exec("_ref_trainer.layers", os.path.join(root, "trainer", "layers.py"), "_ref_trainer")
    ns.reg = _exec("_ref_trainer.reg", os.path.join(root, "trainer", "reg.py"), "_ref_trainer")
    ns.transformer = _exec("_ref_trainer.transformer", os.path.join(root, "trainer", "transformer.py"), "_ref_trainer")
    ns.utils = _exec("_ref_trainer.utils", os.path.join(root, "trainer", "utils.py"), "_ref_trainer")
    return ns
