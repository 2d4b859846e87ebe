import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for path in (ROOT, os.path.join(ROOT, 'tests')):
    if path not in sys.path:
        sys.path.insert(0, path)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run with -m gpu on the B200 box)')
    # the staged reference converts one-element arrays with int() (pyfds/fields.py:571)
    config.addinivalue_line('filterwarnings',
                            'ignore:Conversion of an array with ndim:DeprecationWarning')


def _cuda_devices():
    """CUDA devices the engine sees (0 without a driver, a GPU or the built library)."""
    try:
        from pyfds_b200 import _build, _engine
        if not os.path.exists(_build.LIBRARY):
            return 0
        return int(_engine.load_library().fds_device_count())
    except Exception:       # noqa: BLE001
        return 0


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped, not failed one by one in fds_create
    (a plain `pytest` on the build container would otherwise bury the CPU results). The library is
    never a reason to skip on a GPU box: there a missing build must fail loudly."""
    gpu_items = [item for item in items if 'gpu' in item.keywords]
    if not gpu_items or _cuda_devices() > 0:
        return
    try:
        import torch
        if torch.cuda.is_available():
            return          # a GPU without a loadable engine: let the tests fail
    except Exception:       # noqa: BLE001
        pass
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in gpu_items:
        item.add_marker(skip)


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
