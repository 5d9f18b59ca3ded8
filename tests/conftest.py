import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'slow: builds a full-size oracle (tens of seconds)')


@pytest.fixture(scope='session')
def built_library():
    ''' make sure the generated code and the shared libraries exist '''
    from aircraft_trajectory_optimization_b200.functions import LIB_PATH
    if not os.path.exists(LIB_PATH) or not os.path.exists(os.path.join(ROOT, 'oracle', 'libsxvm.so')):
        import __graft_entry__
        __graft_entry__.build()
    return LIB_PATH
