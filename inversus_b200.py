"""Import alias: makes the package directory `inversus-reinforcement-learning_b200/` (whose name is
not a valid Python identifier) importable as `inversus_b200`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "inversus-reinforcement-learning_b200")
_spec = importlib.util.spec_from_file_location(
    "inversus_b200", os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["inversus_b200"] = _mod
_spec.loader.exec_module(_mod)
