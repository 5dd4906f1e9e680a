"""Import shim: the package directory is named `3d-mot-using-neural-radiance-fields_b200`, which is
not a valid Python identifier, so it is loaded here under the module name `star_b200`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "3d-mot-using-neural-radiance-fields_b200")


def _load():
    spec = importlib.util.spec_from_file_location("star_b200", os.path.join(_DIR, "__init__.py"),
                                                  submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["star_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


_load()
