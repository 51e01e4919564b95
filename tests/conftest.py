import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA GPU (B200)')


@pytest.fixture(scope='session')
def lib():
    """The C-ABI library bound to cuda:0 (GPU tests only)."""
    import __graft_entry__ as g
    g.build()
    from cardoon_b200._lib import get_lib
    return get_lib(0)
