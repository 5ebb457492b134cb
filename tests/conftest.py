import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    from svr2_import import load_package
    return load_package()


@pytest.fixture(scope="session")
def svr2lib(pkg):
    import importlib
    return importlib.import_module("comfyui_seedvr2_videoupscaler_b200.lib")
