import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, 'tests')):
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')


@pytest.fixture(scope='session')
def library():
    """Path of libfdsb200.so, (re)built from source when nvcc is available and it is stale."""
    from pyfds_b200 import _build
    try:
        return _build.build_library()
    except RuntimeError:
        if os.path.exists(_build.LIBRARY):
            return _build.LIBRARY
        raise


def bits(array):
    """View float64 data as int64 so that comparisons distinguish -0.0 from 0.0 and NaN payloads."""
    import numpy as np
    array = np.ascontiguousarray(array)
    return array.view(np.int64) if array.dtype == np.float64 else array
