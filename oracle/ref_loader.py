"""Import the UNMODIFIED reference files on the mini-DGL shim.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Works only where
``/root/reference`` exists, i.e. in the dev container; nothing that runs on the
GPU box may call this.  The reference modules are loaded under private names
(``_kgat_ref_models`` ...) so they never shadow the product's own ``models`` /
``metric`` modules.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

from . import dgl_shim

REFERENCE_DIR = os.environ.get("KGAT_REFERENCE_DIR", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "models.py"))


def _load(name):
    key = "_kgat_ref_" + name
    if key in sys.modules:
        return sys.modules[key]
    if not available():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_DIR)
    dgl_shim.install()
    if not hasattr(np, "asfarray"):          # removed in NumPy 2; metric.py:6,15 use it
        np.asfarray = lambda a, dtype=np.float64: np.asarray(a, dtype=dtype)
    spec = importlib.util.spec_from_file_location(key, os.path.join(REFERENCE_DIR, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[key] = mod
    spec.loader.exec_module(mod)
    return mod


def models():
    return _load("models")


def metric():
    return _load("metric")


def dataset():
    """dataset.py imports fine on the shim; only its pure-numpy helpers are usable
    (``DataLoader.__init__`` needs ``DataFrame.append``, gone in pandas >= 2)."""
    return _load("dataset")
