"""Import shim: makes the package directory ``star-gcn_b200/`` importable as ``stargcn_b200``
(a hyphen cannot appear in a Python module name)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "star-gcn_b200")
_spec = importlib.util.spec_from_file_location("stargcn_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["stargcn_b200"] = _mod
_spec.loader.exec_module(_mod)
