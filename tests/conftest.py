import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def tess():
    """The product package (hyphenated directory name -> importlib)."""
    return importlib.import_module("the-tessellator_b200")


@pytest.fixture(scope="session")
def gen(tess):
    return tess.generators


@pytest.fixture(scope="session")
def ob():
    import oracle_binding

    oracle_binding.lib()
    return oracle_binding
