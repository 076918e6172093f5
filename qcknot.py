"""Import shim: the package directory `quantumcollocation.jl_b200/` has a dot in its name, so it is loaded here
under the module name `qcknot` (`import qcknot` from the repo root)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quantumcollocation.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "qcknot", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["qcknot"] = _mod
_spec.loader.exec_module(_mod)
