"""Import shim: the package directory is named `trafficbotsv1.5_b200` (not a valid Python identifier),
so it is registered under the module name `trafficbotsv1_5_b200`. Usage: `import tbpkg; tb = tbpkg.load()`
or, after `import tbpkg`, `from trafficbotsv1_5_b200 import ...`."""
import importlib.util
import os
import sys

NAME = "trafficbotsv1_5_b200"
ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "trafficbotsv1.5_b200")


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod


load()
