"""Import shim: the package directory is named ``bijectors.jl_b200`` (with a dot), which Python's import
system cannot address directly.  Importing ``bijectors_jl_b200`` loads that directory as a regular package
under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bijectors.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "bijectors_jl_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir]
)
_mod = importlib.util.module_from_spec(_spec)
sys.modules["bijectors_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
