"""Loader for the in-tree sm_100a extension (``baton_b200/_C.so``).

The extension is the product: on a machine with a CUDA device every op in
``baton_b200.ops`` runs its hand-written kernel and a missing/unloadable
extension is a hard error (no silent eager fallback).  On a GPU-less host the
module still imports (``nvcc`` cross-compiles there) but is never called.
"""
from __future__ import annotations

import importlib
import os

_C = None
_ERR = None


def load(build_if_missing: bool = False):
    global _C, _ERR
    if _C is not None:
        return _C
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = os.path.join(here, "_C.so")
    if not os.path.exists(so) and build_if_missing:
        from .. import build_ext
        build_ext.build()
    try:
        _C = importlib.import_module("baton_b200._C")
    except Exception as exc:  # pragma: no cover - exercised only on broken installs
        _ERR = exc
        raise RuntimeError(
            "baton_b200._C (sm_100a kernels) is not available: {!r}. "
            "Run `python -m baton_b200.build_ext`.".format(exc)) from exc
    return _C


def available() -> bool:
    try:
        load()
        return True
    except Exception:
        return False
