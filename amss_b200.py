"""Import shim: exposes the package directory `adaptive-multispeaker-separation_b200/` (whose
name is not a valid Python identifier) as the module `amss_b200`."""
import importlib.util
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "adaptive-multispeaker-separation_b200")
_spec = importlib.util.spec_from_file_location(
    "amss_b200", os.path.join(_PKG, "__init__.py"), submodule_search_locations=[_PKG])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["amss_b200"] = _mod
_spec.loader.exec_module(_mod)
