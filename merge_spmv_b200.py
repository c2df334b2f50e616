"""Import alias: the package directory is named ``merge-spmv_b200`` (not a Python identifier),
so ``import merge_spmv_b200`` loads it from there."""
import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "merge-spmv_b200")
_spec = _ilu.spec_from_file_location(
    "merge_spmv_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = _ilu.module_from_spec(_spec)
_sys.modules["merge_spmv_b200"] = _mod
_spec.loader.exec_module(_mod)
