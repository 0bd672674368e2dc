"""Import shim: the product package lives in the directory `dftk.jl_b200/` (a name Python cannot
import directly); `import dftk_b200` loads it under that module name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dftk.jl_b200")
_spec = importlib.util.spec_from_file_location("dftk_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["dftk_b200"] = _mod
_spec.loader.exec_module(_mod)
