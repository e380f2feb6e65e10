import importlib
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def rtb():
    """The product package (ctypes over the C ABI).  Builds in-tree if the library is missing."""
    mod = importlib.import_module("ray-tracing-v06_b200")
    if not mod.LIB_PATH.exists() or not mod.SCENES_LIB_PATH.exists():
        mod.build()
    mod.lib(); mod.scenes_lib()
    return mod


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    mod = importlib.import_module("pyoracle")
    mod.lib()
    return mod


GOLDEN = ROOT / "tests" / "golden"
