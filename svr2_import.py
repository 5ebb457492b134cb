"""Import shim: the package directory is ``comfyui-seedvr2_videoupscaler_b200/`` (hyphenated, as
the project layout names it), which is not a valid Python identifier.  This
registers it as ``comfyui_seedvr2_videoupscaler_b200`` in ``sys.modules``."""
from __future__ import annotations

import importlib.util
import os
import sys

PKG_NAME = "comfyui_seedvr2_videoupscaler_b200"
PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "comfyui-seedvr2_videoupscaler_b200")


def load_package():
    if PKG_NAME in sys.modules:
        return sys.modules[PKG_NAME]
    spec = importlib.util.spec_from_file_location(
        PKG_NAME, os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[PKG_NAME] = mod
    spec.loader.exec_module(mod)
    return mod
