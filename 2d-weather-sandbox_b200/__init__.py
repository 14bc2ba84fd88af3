"""wsb200 — B200-native simulation core for 2D-Weather-Sandbox's per-iteration loop.

The directory name (`2d-weather-sandbox_b200`) is not a Python identifier; import it as

    import importlib; wsb = importlib.import_module("2d-weather-sandbox_b200")

or through the alias module `wsb200` at the repository root (`import wsb200`).

Only host-side code lives in Python: the save-file codec, the parameter derivation the reference
does in JavaScript, the x-strip plan and a ctypes binding of the C ABI (include/wsb200.h).  All
simulation arithmetic runs in csrc/libwsb200.so (hand-written sm_100a CUDA); there is no CPU
fallback — constructing a `Simulation` without the built library raises.
"""
from . import multi, params, savefile, strips, synth  # noqa: F401
from .sim import Simulation, load_library, library_path  # noqa: F401

__all__ = ["multi", "params", "savefile", "strips", "synth", "Simulation", "load_library", "library_path"]
