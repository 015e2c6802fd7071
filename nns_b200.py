"""Import alias: ``import nns_b200`` loads the package that lives in the (hyphenated, hence not
directly importable) directory ``neural-navier-stokes_b200/``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "neural-navier-stokes_b200")
_spec = importlib.util.spec_from_file_location(
    "nns_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["nns_b200"] = _mod
_spec.loader.exec_module(_mod)
