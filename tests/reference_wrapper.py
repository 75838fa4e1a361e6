"""Loads the reference's unchanged Python wrapper (baseline/_ref/power_grid_model, put there by
tools/install_reference_wrapper.py; in the build container /root/reference/src works too) with libpgm_b200.so as its core
library: the wrapper asks get_pgm_dll_path() where its shared library is, so that one module is answered from here."""
import os
import sys
import types
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "power-grid-model_b200", "libpgm_b200.so")


def wrapper_dir():
    for d in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference/src"):
        if os.path.isfile(os.path.join(d, "power_grid_model", "__init__.py")):
            return d
    return None


def load():
    """returns the imported `power_grid_model` package or None when the wrapper is not available"""
    if "power_grid_model" in sys.modules:
        return sys.modules["power_grid_model"]
    d = wrapper_dir()
    if d is None:
        return None
    name = "power_grid_model._core.power_grid_model_c.get_pgm_dll_path"
    fake = types.ModuleType(name)
    fake.get_pgm_dll_path = lambda: Path(LIB)
    sys.modules[name] = fake
    sys.path.insert(0, d)
    import power_grid_model

    return power_grid_model
