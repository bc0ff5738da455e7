"""Import shim: the package directory `3d-astar-path-planners_b200/` is not a valid Python identifier, so
this module loads it under the importable name `astar_planners_b200` (sub-module `synth` included)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "3d-astar-path-planners_b200")


def _load():
    name = "astar_planners_b200"
    spec = importlib.util.spec_from_file_location(name, os.path.join(_PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_load()
