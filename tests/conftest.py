import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def have_reference_build() -> bool:
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_vamp.so"))


@pytest.fixture(scope="session")
def ctx():
    import ompl_b200
    c = ompl_b200.Context(0)
    yield c
    c.close()
