"""Import shim: the product package lives in the directory `grid.jl_b200/` (the reference's name
plus `_b200`), which is not a valid Python identifier.  `import grid_jl_b200` loads that directory
as a regular package under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "grid.jl_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
