import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    """tests go through the C ABI: make sure the in-tree .so exists (cross-compiles without a GPU)"""
    import __graft_entry__ as g
    g.build()
