import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'jax-cpfem_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests'), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')
    config.addinivalue_line('markers', 'slow: replays of whole committed series (the longest GPU tests; deselect with -m "gpu and not slow")')


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` need a CUDA device AND the built extension: skip them (instead of erroring at import of the
    first plan) on a machine that has neither.  On a GPU box a missing library is a failure, not a skip."""
    try:
        import torch
        have_cuda = torch.cuda.is_available()
    except Exception:
        have_cuda = False
    if have_cuda:
        return
    skip = pytest.mark.skip(reason='needs a CUDA device (B200); run with -m gpu on the GPU box')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def hostcheck():
    """Host build of the product's per-point header (tests/hostcheck) - test infrastructure only."""
    import hostcheck_build
    return hostcheck_build.load()
