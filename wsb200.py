"""Alias: `import wsb200` == the package in ./2d-weather-sandbox_b200 (whose directory name is not
a valid Python identifier)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("2d-weather-sandbox_b200")
sys.modules[__name__] = _pkg
