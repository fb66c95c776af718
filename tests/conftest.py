import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def catalog():
    from mplan2vdl_b200.meta import builtin_catalog
    return builtin_catalog()


def plan_text(name: str) -> str:
    with open(os.path.join(ROOT, "plans", name)) as f:
        return f.read()
