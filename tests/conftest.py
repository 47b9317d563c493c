import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'jax-cpfem_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def hostcheck():
    """Host build of the product's per-point header (tests/hostcheck) - test infrastructure only."""
    import hostcheck_build
    return hostcheck_build.load()
