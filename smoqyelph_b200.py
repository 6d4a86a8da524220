"""Import alias for the package directory `smoqyelphqmc.jl_b200/`.

The directory name mandated for this repo contains a dot, which Python's import system cannot
address directly; this module loads it under the importable name `smoqyelph_b200`.
"""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "smoqyelphqmc.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "smoqyelph_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["smoqyelph_b200"] = _mod
_spec.loader.exec_module(_mod)
