"""Importable alias for the ``t-one_b200`` package (a hyphen cannot appear in an import statement)."""
import importlib as _importlib
import os as _os
import sys as _sys

_root = _os.path.dirname(_os.path.abspath(__file__))
if _root not in _sys.path:
    _sys.path.insert(0, _root)
_pkg = _importlib.import_module("t-one_b200")
_sys.modules[__name__] = _pkg
