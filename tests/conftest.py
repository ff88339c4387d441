import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def ctx_factory(pkg):
    made = []

    def make(m, n, nbr=None):
        c = pkg.Context(0).setup(m, n, nbr)
        made.append(c)
        return c

    yield make
    for c in made:
        c.close()
