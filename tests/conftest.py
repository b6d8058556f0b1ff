import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import __graft_entry__ as g
    o, oc = g.load_oracle()
    oc.build()
    return o, oc


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    so = os.path.join(g.PKG_DIR, "libpls_cuda.so")
    if not os.path.exists(so):
        g.build()
    return g.load_package()


@pytest.fixture(scope="session")
def ctx(pkg):
    c = pkg.Context(0)     # fails loudly (PLS_ECUDA) without a B200: no CPU fallback
    yield c
    c.close()
