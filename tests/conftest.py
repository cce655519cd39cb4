import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

# the reference's nestedCluster hard-codes 3 OpenMP threads: keep libgomp serial
os.environ.setdefault("OMP_THREAD_LIMIT", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run by the driver with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref built from /root/reference")


@pytest.fixture(scope="session")
def oracle():
    from oracle_lib import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle_lib import RefLib

    if not RefLib.available():
        pytest.skip("oracle/_ref/libklsh_ref.so not built (needs /root/reference)")
    return RefLib()


@pytest.fixture(scope="session")
def gpu():
    """A klsh context on cuda:0.  No fallback: a missing library or GPU is a failure."""
    from kmerlsh_b200 import Context

    ctx = Context(0)
    yield ctx
    ctx.close()
