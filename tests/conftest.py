import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu)")


def _ensure_built():
    so = os.path.join(ROOT, "candlezip_b200", "libcandlezip_b200.so")
    if not os.path.exists(so):
        import __graft_entry__ as g

        g.build()


_ensure_built()


@pytest.fixture(scope="session")
def fixtures():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "reference_fixtures.npz"))


@pytest.fixture(scope="session")
def gpu_ctx():
    """A real device context.  GPU tests must run the CUDA path: no device -> the test FAILS (never skips to a fallback)."""
    import candlezip_b200 as cz

    ctx = cz.Context(0)
    yield ctx
    ctx.close()
