import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _load_pkg():
    if "meagre_crowd_b200" in sys.modules:
        return sys.modules["meagre_crowd_b200"]
    d = os.path.join(ROOT, "meagre-crowd_b200")
    spec = importlib.util.spec_from_file_location("meagre_crowd_b200", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    m = importlib.util.module_from_spec(spec)
    sys.modules["meagre_crowd_b200"] = m
    spec.loader.exec_module(m)
    if not os.path.exists(m.LIB_PATH) or not os.path.exists(m.CLI_PATH):
        m.build()
    return m


_load_pkg()   # registers module `meagre_crowd_b200` so test modules can import its submodules


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def mc():
    return _load_pkg()


@pytest.fixture(scope="session")
def golden():
    return GOLDEN
