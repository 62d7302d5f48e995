import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree libraries exist (no-op when they are up to date)."""
    import __graft_entry__ as g
    need = [os.path.join(ROOT, "gmx2020postanalysis_b200", "libb2a.so"), os.path.join(ROOT, "workloads", "libsynthgen.so")]
    need.append(os.path.join(ROOT, "oracle", "liboracle.so"))
    if not all(os.path.exists(n) for n in need):
        g.build()


@pytest.fixture(scope="session")
def gpu_ctx_factory():
    from gmx2020postanalysis_b200 import lib
    if lib.load().b2a_device_count() == 0:
        pytest.fail("gpu test selected but no CUDA device is visible (there is no CPU fallback)")
    made = []

    def make(natoms, K, device=0):
        c = lib.Context(natoms, K, device)
        made.append(c)
        return c

    yield make
    for c in made:
        c.close()
