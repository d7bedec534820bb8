"""Importable alias of the package directory `causal-domain-clustering-for-multi-domain-recommendation_b200/`
(its name is not a valid Python identifier):  `import cdcmdr_b200 as cm; cm.PLE(...)`."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "causal-domain-clustering-for-multi-domain-recommendation_b200")
_NAME = "cdcmdr_b200_pkg"


def _load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                  submodule_search_locations=[_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


_pkg = _load()
globals().update({k: getattr(_pkg, k) for k in dir(_pkg) if not k.startswith("__")})
pkg = _pkg
