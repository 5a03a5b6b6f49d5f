"""Loader for the hyphen-named package directory `rust-ray-tracing-in-a-weekend_b200/`.

The contract fixes the directory name; hyphens are not importable, so this registers it in
sys.modules under the alias `rtw_b200`.  Usage: `import rtw_pkg; rtw_b200 = rtw_pkg.load()`.
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "rust-ray-tracing-in-a-weekend_b200")
ALIAS = "rtw_b200"


def load():
    if ALIAS in sys.modules:
        return sys.modules[ALIAS]
    spec = importlib.util.spec_from_file_location(
        ALIAS, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[ALIAS] = mod
    spec.loader.exec_module(mod)
    return mod
