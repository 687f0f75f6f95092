"""Importable alias of the `big-linear-algebra_b200` package (its directory name carries the
reference's hyphens, which the `import` statement cannot spell)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("big-linear-algebra_b200")
globals().update({k: v for k, v in vars(_pkg).items() if not k.startswith("__")})
